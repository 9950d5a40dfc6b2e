import sys; sys.path.insert(0, "."); sys.path.insert(0, "tests/golden"); sys.path.insert(0, "tests")
import numpy as np
import train_cases as TC
import flownet2_b200 as F
from oracle.net import OracleNet, synth_weights
gold = np.load("tests/golden/train_golden.npz")
proto, ins = TC.loss_net_proto(), TC.loss_net_inputs()
small = F.fill_template(F.model_template("FlowNet2-C"), 64, 64)
_, cm = synth_weights(small, TC.LOSS_NET["seed"], F.fill_template(F.model_template("FlowNet2-C"), 192, 100))
on = OracleNet(proto, cm, batch=1, f64acc=True); B = on.forward(**ins); _, P = on.backward()
net = F.Net(proto, cm, F.TEST, batch=1); out = net.forward(**ins); net.clear_param_diffs(); net.backward()
worst_o = worst_e = 0
for k in [k for k in gold.files if k.startswith("N/lossnet/grad/")]:
    _, _, _, name, i = k.split("/"); want = gold[k]; sc = float(np.abs(want).max())
    eo = float(np.abs(TC.grad_signature(name, int(i), P[name][int(i)]) - want).max()) / sc
    ee = float(np.abs(TC.grad_signature(name, int(i), net.param(name, int(i), diff=True)) - want).max()) / sc
    worst_o = max(worst_o, eo); worst_e = max(worst_e, ee)
    if max(eo, ee) > 2e-5: print("%-28s oracle-ref %.2e engine-ref %.2e" % (k[15:], eo, ee))
print("worst: oracle vs reference %.2e, engine vs reference %.2e" % (worst_o, worst_e))
for l in TC.LOSS_NET["weights"]:
    print(l, float(gold["N/lossnet/loss%d" % l][0]), float(B["flow_loss%d" % l][0]), float(out["flow_loss%d" % l].reshape(-1)[0]))
