"""Per-blob / per-parameter gradient error of Net::Backward against OracleNet.backward (diagnostic)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import flownet2_b200 as fn2
from oracle.net import OracleNet
from tests.util import maxabs, rng, smooth_images

SEEDS = ["predict_flow6", "predict_flow5", "predict_flow4", "predict_flow3", "predict_flow2"]
w, h, batch = 192, 100, 1
proto = fn2.fill_template(fn2.model_template("FlowNet2-C"), w, h)
net = fn2.Net(proto, None, fn2.TEST, batch=batch)
net.fill_params(7)
weights = net.to_caffemodel()
img0, img1 = smooth_images(rng(7), batch, h, w)
net.forward(img0=img0, img1=img1)
r = rng(8)
seeds = {s: r.standard_normal(net.blobs[s].shape).astype(np.float32) for s in SEEDS}
net.clear_param_diffs()
net.backward(**seeds)
onet = OracleNet(proto, weights, batch=batch, f64acc=True)
B = onet.forward(img0=img0, img1=img1)
D, P = onet.backward(**seeds)
def rel(a, b):
    return maxabs(a, b) / max(float(np.abs(b).max()), 1e-30)
for k in D:
    if k in SEEDS: continue
    g = net.get_diff(k)
    e = np.abs(g.astype(np.float64) - D[k])
    idx = np.unravel_index(e.argmax(), e.shape)
    print("blob %-22s rel %.3e  absmax %.3e  worst@%s got %.6g want %.6g  fwd %.6g (oracle %.6g)" % (k, rel(g, D[k]), np.abs(D[k]).max(), idx, g[idx], D[k][idx],
          net.blobs[k].data[idx], B[k][idx]))
for l, gs in P.items():
    for i, g in enumerate(gs):
        print("param %-20s[%d] rel %.3e" % (l, i, rel(net.param(l, i, diff=True).reshape(g.shape), g)))
