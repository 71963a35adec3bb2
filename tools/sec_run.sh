HTV_DEBUG=1 python tools/run_one.py l 13500000 1 64 2>&1 | tail -4
HTV_DEBUG=1 python tools/run_one.py l 16000000 1 64 2>&1 | tail -2
python -m pytest tests -m gpu -x -q -k "secam or long or parity or chunk or vbi or dropin" 2>&1 | tail -3
python tools/sec_time.py
