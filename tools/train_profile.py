"""Per-layer device times of one FlowNet2-C training step (forward: fn2_net_time_layers; backward: FN2_BWD_PROFILE=1)."""
import os, sys
os.environ["FN2_BWD_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import flownet2_b200 as F
proto = F.fill_train_template(F.train_template("FlowNet2-C"), 448, 320, 512, 384, 8)
net = F.Net(proto, None, F.TRAIN); net.fill_params(1)
r = np.random.default_rng(0)
ins = dict(img0=r.uniform(0, 255, (8, 3, 384, 512)).astype(np.float32), img1=r.uniform(0, 255, (8, 3, 384, 512)).astype(np.float32),
           flow_gt=r.standard_normal((8, 2, 384, 512)).astype(np.float32))
for it in range(2):
    net.forward(**ins); net.clear_param_diffs()
    if it == 1: sys.stderr.write("---- second pass ----\n")
    net.backward(); net.sync()
lt = net.time_layers()
for n, t, ms in sorted(lt, key=lambda x: -x[2])[:12]: print("fwd %-22s %-16s %.3f" % (n, t, ms))
print("fwd total", sum(x[2] for x in lt))
