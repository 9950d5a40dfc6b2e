// extern "C" net-level entry points of include/fn2.h over caffe::Net<float>.
#include <cstring>

#include "net.hpp"
#include "hdf5_min.hpp"

namespace fn2 { void set_error(const char* fmt, ...); }

struct fn2_net {
    std::unique_ptr<caffe::Net<float> > net;
    std::vector<std::string> input_names, output_names;
    std::string caffemodel_cache;
};

#define FN2_TRY try {
#define FN2_CATCH                                                        \
    } catch (const caffe::ParseError& e) {                               \
        fn2::set_error("%s", e.what());                                  \
        return FN2_ERR_PARSE;                                            \
    } catch (const std::exception& e) {                                  \
        fn2::set_error("%s", e.what());                                  \
        return strstr(e.what(), "Unknown") ? FN2_ERR_NOTFOUND : FN2_ERR_INVALID; \
    }

namespace caffe { void SampleAugmentationCoeffs(const LayerParameter& lp, uint32_t seed, int num, int width, int height, float num_iter, float* out); }

extern "C" {

int fn2_net_create(const char* prototxt_text, int phase, fn2_net** out) {
    return fn2_net_create_batch(prototxt_text, phase, 0, out);
}

int fn2_net_create_batch(const char* prototxt_text, int phase, int batch, fn2_net** out) {
    if (!prototxt_text || !out) { fn2::set_error("net_create: null argument"); return FN2_ERR_INVALID; }
    *out = nullptr;
    FN2_TRY
        caffe::NetParameter np = caffe::NetParameter::FromText(prototxt_text);
        if (batch > 0) {
            for (auto& lp : np.layers) {
                if (lp.type() != "Input") continue;
                caffe::Message* ip = lp.m->mutable_msg("input_param");
                for (auto& f : ip->fields)
                    if (f.name == "shape" && f.is_msg)
                        for (auto& d : f.msg->fields)
                            if (d.name == "dim") { d.scalar = std::to_string(batch); break; }
            }
        }
        std::unique_ptr<fn2_net> h(new fn2_net());
        h->net.reset(new caffe::Net<float>(np, phase == 0 ? caffe::TRAIN : caffe::TEST));
        for (int i : h->net->input_blob_indices()) h->input_names.push_back(h->net->blob_names()[i]);
        for (int i : h->net->output_blob_indices()) h->output_names.push_back(h->net->blob_names()[i]);
        *out = h.release();
        return FN2_OK;
    FN2_CATCH
}

void fn2_net_destroy(fn2_net* net) { delete net; }

int fn2_net_copy_trained_layers(fn2_net* net, const void* caffemodel, size_t bytes) {
    if (!net || !caffemodel) { fn2::set_error("copy_trained_layers: null argument"); return FN2_ERR_INVALID; }
    FN2_TRY
        net->net->CopyTrainedLayersFrom(caffemodel, bytes);
        return FN2_OK;
    FN2_CATCH
}

int fn2_net_to_caffemodel(fn2_net* net, void* buf, size_t* bytes) {
    if (!net || !bytes) { fn2::set_error("to_caffemodel: null argument"); return FN2_ERR_INVALID; }
    FN2_TRY
        if (!buf || net->caffemodel_cache.empty()) net->caffemodel_cache = net->net->ToCaffemodel();
        if (!buf) { *bytes = net->caffemodel_cache.size(); return FN2_OK; }
        if (*bytes < net->caffemodel_cache.size()) { fn2::set_error("to_caffemodel: buffer too small"); return FN2_ERR_INVALID; }
        memcpy(buf, net->caffemodel_cache.data(), net->caffemodel_cache.size());
        *bytes = net->caffemodel_cache.size();
        net->caffemodel_cache.clear();
        net->caffemodel_cache.shrink_to_fit();
        return FN2_OK;
    FN2_CATCH
}

int fn2_net_fill_params(fn2_net* net, uint64_t seed) {
    if (!net) { fn2::set_error("fill_params: null net"); return FN2_ERR_INVALID; }
    FN2_TRY
        net->net->FillParams(seed);
        return FN2_OK;
    FN2_CATCH
}

int fn2_net_param_arena(fn2_net* net, void** dev_ptr, size_t* bytes) {
    if (!net || !dev_ptr || !bytes) { fn2::set_error("param_arena: null argument"); return FN2_ERR_INVALID; }
    net->net->ParamArena(dev_ptr, bytes);
    return FN2_OK;
}

int fn2_net_params_changed(fn2_net* net) {
    if (!net) { fn2::set_error("params_changed: null net"); return FN2_ERR_INVALID; }
    FN2_TRY
        net->net->ArenaWritten();      // the caller wrote the DEVICE arena: host caches of every param blob are stale
        return FN2_OK;
    FN2_CATCH
}

int fn2_net_num_inputs(fn2_net* net) { return net ? (int)net->input_names.size() : 0; }
const char* fn2_net_input_name(fn2_net* net, int i) {
    return (net && i >= 0 && i < (int)net->input_names.size()) ? net->input_names[i].c_str() : nullptr;
}
int fn2_net_num_outputs(fn2_net* net) { return net ? (int)net->output_names.size() : 0; }
const char* fn2_net_output_name(fn2_net* net, int i) {
    return (net && i >= 0 && i < (int)net->output_names.size()) ? net->output_names[i].c_str() : nullptr;
}
int fn2_net_num_blobs(fn2_net* net) { return net ? (int)net->net->blob_names().size() : 0; }
const char* fn2_net_blob_name(fn2_net* net, int i) {
    return (net && i >= 0 && i < (int)net->net->blob_names().size()) ? net->net->blob_names()[i].c_str() : nullptr;
}
int fn2_net_num_layers(fn2_net* net) { return net ? (int)net->net->layer_names().size() : 0; }
const char* fn2_net_layer_name(fn2_net* net, int i) {
    return (net && i >= 0 && i < (int)net->net->layer_names().size()) ? net->net->layer_names()[i].c_str() : nullptr;
}
const char* fn2_net_layer_type(fn2_net* net, int i) {
    return (net && i >= 0 && i < (int)net->net->layers().size()) ? net->net->layers()[i]->type() : nullptr;
}

int fn2_net_blob_shape(fn2_net* net, const char* blob, int shape[4]) {
    if (!net || !blob || !shape) { fn2::set_error("blob_shape: null argument"); return FN2_ERR_INVALID; }
    FN2_TRY
        auto b = net->net->blob_by_name(blob);
        for (int i = 0; i < 4; i++) shape[i] = b->shape(i);
        return FN2_OK;
    FN2_CATCH
}

int fn2_net_set_input(fn2_net* net, const char* blob, const float* host_nchw) {
    if (!net || !blob || !host_nchw) { fn2::set_error("set_input: null argument"); return FN2_ERR_INVALID; }
    FN2_TRY
        net->net->SetInput(blob, host_nchw);
        return FN2_OK;
    FN2_CATCH
}
int fn2_net_set_input_device(fn2_net* net, const char* blob, const float* dev_nchw) {
    if (!net || !blob || !dev_nchw) { fn2::set_error("set_input_device: null argument"); return FN2_ERR_INVALID; }
    FN2_TRY
        net->net->SetInputDevice(blob, dev_nchw);
        return FN2_OK;
    FN2_CATCH
}
int fn2_net_get_blob(fn2_net* net, const char* blob, float* host_nchw) {
    if (!net || !blob || !host_nchw) { fn2::set_error("get_blob: null argument"); return FN2_ERR_INVALID; }
    FN2_TRY
        net->net->GetBlob(blob, host_nchw);
        return FN2_OK;
    FN2_CATCH
}
int fn2_net_get_blob_device(fn2_net* net, const char* blob, float* dev_nchw) {
    if (!net || !blob || !dev_nchw) { fn2::set_error("get_blob_device: null argument"); return FN2_ERR_INVALID; }
    FN2_TRY
        net->net->GetBlobDevice(blob, dev_nchw);
        return FN2_OK;
    FN2_CATCH
}

int fn2_net_forward(fn2_net* net) {
    if (!net) { fn2::set_error("forward: null net"); return FN2_ERR_INVALID; }
    FN2_TRY
        net->net->Forward();
        return FN2_OK;
    FN2_CATCH
}
int fn2_net_backward(fn2_net* net) {
    if (!net) { fn2::set_error("backward: null net"); return FN2_ERR_INVALID; }
    FN2_TRY
        net->net->Backward();
        return FN2_OK;
    FN2_CATCH
}
int fn2_net_clear_param_diffs(fn2_net* net) {
    if (!net) { fn2::set_error("clear_param_diffs: null net"); return FN2_ERR_INVALID; }
    FN2_TRY
        net->net->ClearParamDiffs();
        return FN2_OK;
    FN2_CATCH
}
int fn2_net_set_diff(fn2_net* net, const char* blob, const float* host_nchw) {
    if (!net || !blob || !host_nchw) { fn2::set_error("set_diff: null argument"); return FN2_ERR_INVALID; }
    FN2_TRY
        net->net->SetDiff(blob, host_nchw);
        return FN2_OK;
    FN2_CATCH
}
int fn2_net_get_diff(fn2_net* net, const char* blob, float* host_nchw) {
    if (!net || !blob || !host_nchw) { fn2::set_error("get_diff: null argument"); return FN2_ERR_INVALID; }
    FN2_TRY
        net->net->GetDiff(blob, host_nchw);
        return FN2_OK;
    FN2_CATCH
}
int fn2_net_param_diff_arena(fn2_net* net, void** dev_ptr, size_t* bytes) {
    if (!net || !dev_ptr || !bytes) { fn2::set_error("param_diff_arena: null argument"); return FN2_ERR_INVALID; }
    FN2_TRY
        net->net->ParamDiffArena(dev_ptr, bytes);
        return FN2_OK;
    FN2_CATCH
}
static caffe::Blob<float>* find_param(fn2_net* net, const char* layer, int index) {
    const auto& names = net->net->layer_names();
    for (size_t i = 0; i < names.size(); i++)
        if (names[i] == layer) {
            auto& bl = net->net->layers()[i]->blobs();
            if (index < 0 || index >= (int)bl.size()) return nullptr;
            return bl[index].get();
        }
    return nullptr;
}
int fn2_net_param_shape(fn2_net* net, const char* layer, int index, int shape[4]) {
    if (!net || !layer || !shape) { fn2::set_error("param_shape: null argument"); return FN2_ERR_INVALID; }
    FN2_TRY
        caffe::Blob<float>* b = find_param(net, layer, index);
        if (!b) { fn2::set_error("param_shape: no parameter blob %s[%d]", layer, index); return FN2_ERR_INVALID; }
        for (int i = 0; i < 4; i++) shape[i] = b->shape(i);
        return FN2_OK;
    FN2_CATCH
}
int fn2_net_get_param(fn2_net* net, const char* layer, int index, int diff, float* host) {
    if (!net || !layer || !host) { fn2::set_error("get_param: null argument"); return FN2_ERR_INVALID; }
    FN2_TRY
        caffe::Blob<float>* b = find_param(net, layer, index);
        if (!b) { fn2::set_error("get_param: no parameter blob %s[%d]", layer, index); return FN2_ERR_INVALID; }
        caffe::Caffe::stream() = net->net->stream();
        net->net->Sync();
        if (diff) { void* a; size_t n; net->net->ParamDiffArena(&a, &n); }
        const float* p = diff ? b->cpu_diff() : b->cpu_data();
        memcpy(host, p, (size_t)b->count() * sizeof(float));
        return FN2_OK;
    FN2_CATCH
}
int fn2_net_launches_per_backward(fn2_net* net) { return net ? net->net->launches_per_backward() : 0; }
int fn2_net_layer_need_backward(fn2_net* net, int layer) {
    if (!net || layer < 0 || layer >= (int)net->net->layers().size()) return -1;
    FN2_TRY
        return net->net->layer_need_backward()[layer] ? 1 : 0;
    } catch (...) { return -1; }
}

int fn2_net_sync(fn2_net* net) {
    if (!net) { fn2::set_error("sync: null net"); return FN2_ERR_INVALID; }
    FN2_TRY
        net->net->Sync();
        return FN2_OK;
    FN2_CATCH
}
void* fn2_net_stream(fn2_net* net) { return net ? (void*)net->net->stream() : nullptr; }

int fn2_net_time_layers(fn2_net* net, float* ms) {
    if (!net || !ms) { fn2::set_error("time_layers: null argument"); return FN2_ERR_INVALID; }
    FN2_TRY
        net->net->TimeLayers(ms);
        return FN2_OK;
    FN2_CATCH
}
int fn2_net_launches_per_forward(fn2_net* net) { return net ? net->net->launches_per_forward() : 0; }
int fn2_net_graph_active(fn2_net* net) { return net && net->net->graph_active() ? 1 : 0; }

int fn2_net_layer_work(fn2_net* net, int layer, double* flops, double* bytes) {
    if (!net || !flops || !bytes || layer < 0 || layer >= (int)net->net->layers().size()) {
        fn2::set_error("layer_work: bad argument");
        return FN2_ERR_INVALID;
    }
    net->net->LayerWork(layer, flops, bytes);
    return FN2_OK;
}

static int emit_string(const std::string& s, char* out, size_t* bytes) {
    if (!bytes) { fn2::set_error("null size pointer"); return FN2_ERR_INVALID; }
    if (!out) { *bytes = s.size() + 1; return FN2_OK; }
    if (*bytes < s.size() + 1) { fn2::set_error("buffer too small"); return FN2_ERR_INVALID; }
    memcpy(out, s.c_str(), s.size() + 1);
    *bytes = s.size() + 1;
    return FN2_OK;
}

int fn2_proto_canonical(const char* prototxt_text, char* out, size_t* bytes) {
    if (!prototxt_text) { fn2::set_error("proto_canonical: null text"); return FN2_ERR_INVALID; }
    FN2_TRY
        caffe::NetParameter np = caffe::NetParameter::FromText(prototxt_text);
        std::string s;
        if (!np.name.empty()) s += "name: \"" + np.name + "\"\n";
        for (const auto& l : np.layers) s += "layer {\n" + caffe::PrintTextFormat(*l.m, 1) + "}\n";
        return emit_string(s, out, bytes);
    FN2_CATCH
}

int fn2_aug_sample(const char* layer_prototxt, unsigned int seed, int num, int width, int height, float num_iter, float* coeffs_out) {
    if (!layer_prototxt || !coeffs_out || num <= 0) { fn2::set_error("aug_sample: invalid argument"); return FN2_ERR_INVALID; }
    FN2_TRY
        caffe::NetParameter np = caffe::NetParameter::FromText(layer_prototxt);
        const caffe::LayerParameter* lp = nullptr;
        for (const auto& l : np.layers) if (l.type() == "DataAugmentation") { lp = &l; break; }
        if (!lp) { fn2::set_error("aug_sample: no DataAugmentation layer in the text"); return FN2_ERR_INVALID; }
        caffe::SampleAugmentationCoeffs(*lp, seed, num, width, height, num_iter, coeffs_out);
        return FN2_OK;
    FN2_CATCH
}

static std::string caffemodel_to_h5(const void* caffemodel, size_t n) {
    std::vector<caffe::LayerBlobs> ls = caffe::ParseCaffemodel(caffemodel, n);
    std::vector<std::pair<std::string, std::vector<caffe::H5Blob> > > layers;
    for (const auto& l : ls) {
        if (l.blobs.empty()) continue;
        std::vector<caffe::H5Blob> bl;
        for (const auto& b : l.blobs) { caffe::H5Blob hb; hb.dims = b.shape; hb.data = b.data; bl.push_back(std::move(hb)); }
        layers.push_back({l.name, std::move(bl)});
    }
    return caffe::WriteCaffemodelH5(layers);
}

int fn2_caffemodel_to_hdf5(const void* caffemodel, size_t n, void* out, size_t* bytes) {
    if (!caffemodel || !bytes) { fn2::set_error("caffemodel_to_hdf5: null argument"); return FN2_ERR_INVALID; }
    try {
        const std::string h5 = caffemodel_to_h5(caffemodel, n);
        if (!out) { *bytes = h5.size(); return FN2_OK; }
        if (*bytes < h5.size()) { fn2::set_error("caffemodel_to_hdf5: buffer too small"); return FN2_ERR_INVALID; }
        memcpy(out, h5.data(), h5.size());
        *bytes = h5.size();
        return FN2_OK;
    } catch (const std::exception& e) {
        fn2::set_error("%s", e.what());
        return FN2_ERR_PARSE;
    }
}

int fn2_net_to_hdf5(fn2_net* net, void* buf, size_t* bytes) {
    if (!net || !bytes) { fn2::set_error("to_hdf5: null argument"); return FN2_ERR_INVALID; }
    FN2_TRY
        const std::string cm = net->net->ToCaffemodel();
        const std::string h5 = caffemodel_to_h5(cm.data(), cm.size());
        if (!buf) { *bytes = h5.size(); return FN2_OK; }
        if (*bytes < h5.size()) { fn2::set_error("to_hdf5: buffer too small"); return FN2_ERR_INVALID; }
        memcpy(buf, h5.data(), h5.size());
        *bytes = h5.size();
        return FN2_OK;
    FN2_CATCH
}

int fn2_hdf5_summary(const void* h5, size_t n, char* out, size_t* bytes) {
    if (!h5) { fn2::set_error("hdf5_summary: null data"); return FN2_ERR_INVALID; }
    try {
        caffe::H5File f(h5, n);
        std::string s;
        char buf[128];
        for (const auto& kv : f.datasets()) {
            s += kv.first + " f" + std::to_string(kv.second.elem_size * 8) + " [";
            for (size_t i = 0; i < kv.second.dims.size(); i++) { snprintf(buf, sizeof(buf), i ? ",%d" : "%d", kv.second.dims[i]); s += buf; }
            std::vector<float> v = f.read(kv.first);
            double sum = 0;
            for (float x : v) sum += (double)x;
            snprintf(buf, sizeof(buf), "] n=%zu sum=%.9g first=%.9g last=%.9g", v.size(), sum, v.empty() ? 0.0 : (double)v.front(), v.empty() ? 0.0 : (double)v.back());
            s += buf;
            s += "\n";
        }
        return emit_string(s, out, bytes);
    } catch (const std::exception& e) {
        fn2::set_error("%s", e.what());
        return FN2_ERR_PARSE;
    }
}

int fn2_caffemodel_summary(const void* caffemodel, size_t n, char* out, size_t* bytes) {
    if (!caffemodel) { fn2::set_error("caffemodel_summary: null data"); return FN2_ERR_INVALID; }
    FN2_TRY
        std::vector<caffe::LayerBlobs> ls = caffe::ParseCaffemodel(caffemodel, n);
        std::string s;
        char buf[128];
        for (const auto& l : ls) {
            s += l.name + " " + l.type;
            for (const auto& b : l.blobs) {
                s += " [";
                for (size_t i = 0; i < b.shape.size(); i++) { snprintf(buf, sizeof(buf), i ? ",%d" : "%d", b.shape[i]); s += buf; }
                double sum = 0;
                for (float v : b.data) sum += (double)v;
                snprintf(buf, sizeof(buf), "] n=%zu sum=%.9g", b.data.size(), sum);
                s += buf;
            }
            s += "\n";
        }
        return emit_string(s, out, bytes);
    FN2_CATCH
}

}  // extern "C"
