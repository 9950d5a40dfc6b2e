import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import flownet2_b200 as fn2
from tests.util import maxabs, rng, smooth_images
proto = fn2.fill_template(fn2.model_template("FlowNet2-S"), 64, 64)
a = fn2.Net(proto, None, fn2.TEST)
a.fill_params(5)
blob = a.to_caffemodel()
b = fn2.Net(proto, blob, fn2.TEST)
blob_b = b.to_caffemodel()
print('caffemodel bytes equal:', blob == blob_b, len(blob), len(blob_b))
img0, img1 = smooth_images(rng(5), 1, 64, 64)
a.forward(img0=img0, img1=img1); b.forward(img0=img0, img1=img1)
for name in a.blobs:
    d = maxabs(a.blobs[name].data, b.blobs[name].data)
    if d > 0: print(name, d, a.blobs[name].data.shape)
a2 = fn2.Net(proto, None, fn2.TEST); a2.fill_params(5); a2.forward(img0=img0, img1=img1)
print("a vs fresh a2 (no to_caffemodel before forward):", maxabs(a.blobs["predict_flow_final"].data, a2.blobs["predict_flow_final"].data))
b2 = fn2.Net(proto, blob, fn2.TEST); b2.forward(img0=img0, img1=img1)
print("b vs b2:", maxabs(b.blobs["predict_flow_final"].data, b2.blobs["predict_flow_final"].data))
