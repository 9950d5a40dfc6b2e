"""Two FlowNet2-C training steps at the config 5 shape (for ncu launch lists: profile the second one)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import flownet2_b200 as F
proto = F.fill_train_template(F.train_template("FlowNet2-C"), 448, 320, 512, 384, 8)
net = F.Net(proto, None, F.TRAIN); net.fill_params(1)
r = np.random.default_rng(0)
ins = dict(img0=r.uniform(0, 255, (8, 3, 384, 512)).astype(np.float32), img1=r.uniform(0, 255, (8, 3, 384, 512)).astype(np.float32),
           flow_gt=r.standard_normal((8, 2, 384, 512)).astype(np.float32))
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    net.forward(**ins); net.clear_param_diffs(); net.backward(); net.sync()
print("done")
