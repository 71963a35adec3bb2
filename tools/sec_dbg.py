import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, hacktv_b200 as H, orc
conf = H.mode_config("l", vfilter=False, noaudio=True)
enc = H.Encoder(conf, 16000000); enc.open_test_source()
try:
    got = enc.render_host(700)
except Exception as e:
    print("ERR", e); got=None
o = orc.Oracle(conf, 16000000); o.open_test_source(); want = o.render(700)
if got is not None:
    d = np.abs(got.astype(np.int32)-want.astype(np.int32)).reshape(700,1024,2)[:,:,0]
    bad = np.argwhere(d>0)
    print("mismatch count", len(bad), "max", d.max())
    if len(bad):
        lines = sorted(set(bad[:,0].tolist()))
        print("first bad lines", lines[:20])
        l=lines[0]; xs=bad[bad[:,0]==l][:,1]; print("line", l, "x range", xs.min(), xs.max(), "n", len(xs))
        g=got.reshape(700,1024,2)[l,:,0]; w=want.reshape(700,1024,2)[l,:,0]
        print(g[xs[:10]], w[xs[:10]])
