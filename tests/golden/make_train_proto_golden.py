"""Parses the authored FlowNet2-C training prototxt (models/FlowNet2-C_train.prototxt.template) with the REFERENCE's own
caffe_pb2 (python/caffe/proto/caffe_pb2.py: any field that is not in the reference's caffe.proto makes text_format.Merge raise) and
writes what it read to tests/golden/ref_pb2_train_prototxt.json, against which the engine's parser is checked.
Run in the container that has /root/reference."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference/python/caffe/proto")


def main():
    import caffe_pb2
    from google.protobuf import text_format
    import flownet2_b200 as F
    txt = F.fill_train_template(F.train_template("FlowNet2-C"), 448, 320, 512, 384, 8)
    net = caffe_pb2.NetParameter()
    text_format.Merge(txt, net)
    layers = []
    for l in net.layer:
        e = {"name": l.name, "type": l.type, "bottom": list(l.bottom), "top": list(l.top), "loss_weight": [float(x) for x in l.loss_weight],
             "propagate_down": [bool(x) for x in l.propagate_down]}
        if l.type == "L1Loss":
            p = l.l1_loss_param
            e["l1"] = [p.l2_per_location, p.l2_prescale_by_channels, p.normalize_by_num_entries, float(p.epsilon), float(p.plateau)]
        if l.HasField("augmentation_param"):
            a = l.augmentation_param
            gens = {}
            for f, v in a.ListFields():
                if f.message_type is not None and f.message_type.name == "RandomGeneratorParameter":
                    gens[f.name] = [v.rand_type, bool(v.exp), float(v.mean), float(v.spread), float(v.prob)]
            e["aug"] = {"crop": [a.crop_width, a.crop_height], "mode": a.mode, "recompute_mean": a.recompute_mean, "generators": gens,
                        "eigvec": [float(x) for x in a.chromatic_eigvec]}
        if l.HasField("coeff_schedule_param"):
            c = l.coeff_schedule_param
            e["schedule"] = [float(c.half_life), float(c.initial_coeff), float(c.final_coeff)]
        if l.type == "Input":
            e["shapes"] = [list(s.dim) for s in l.input_param.shape]
        layers.append(e)
    with open(os.path.join(HERE, "ref_pb2_train_prototxt.json"), "w") as f:
        json.dump({"name": net.name, "layers": layers}, f)
    print("wrote", len(layers), "layers")


if __name__ == "__main__":
    main()
