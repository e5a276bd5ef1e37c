// weights.cu -- state-dict ingestion: PyTorch-layout FP32 host tensors -> folded, repacked device weights.
// Key names follow easyocr/craft.py + easyocr/model/modules.py (detector) and easyocr/model/vgg_model.py (recogniser),
// with the DataParallel "module." prefix already stripped by the caller (easyocr.detection.copyStateDict).
#include "engine.h"

namespace bbocr {

namespace {

struct Dict {
    std::map<std::string, const bbocr_tensor*> m;
    Dict(const bbocr_tensor* t, int n) {
        for (int i = 0; i < n; ++i) {
            std::string k = t[i].name ? t[i].name : "";
            if (k.rfind("module.", 0) == 0) k = k.substr(7);
            m[k] = &t[i];
        }
    }
    const bbocr_tensor* get(const std::string& k, bool required = true) const {
        auto it = m.find(k);
        if (it == m.end()) {
            if (required) fail(BBOCR_E_ARG, "missing tensor '%s' in state dict", k.c_str());
            return nullptr;
        }
        return it->second;
    }
};

int64_t numel(const bbocr_tensor* t) {
    int64_t n = 1;
    for (int i = 0; i < t->ndim; ++i) n *= t->shape[i];
    return n;
}

template <typename T>
T* to_device(Handle* h, const std::vector<T>& v) {
    T* d = nullptr;
    CUDA_CHECK(cudaMalloc(&d, std::max<size_t>(v.size() * sizeof(T), 16)));
    CUDA_CHECK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    h->owned.push_back(d);
    return d;
}

// conv (+ optional BatchNorm, eval mode, eps 1e-5) -> ConvW.  y = scale * conv(x) + bias.
ConvW make_conv(Handle* h, const Dict& d, const std::string& conv, const std::string& bn, int pad, int dil, bool split = false) {
    const bbocr_tensor* w = d.get(conv + ".weight");
    ARG_CHECK(w->ndim == 4, "%s.weight must be 4-D", conv.c_str());
    const bbocr_tensor* b = d.get(conv + ".bias", false);
    ConvW c;
    c.cout = (int)w->shape[0]; c.cin = (int)w->shape[1]; c.kh = (int)w->shape[2]; c.kw = (int)w->shape[3];
    c.pad = pad; c.dil = dil;
    c.cout_pad = (c.cout + 15) / 16 * 16;
    const int taps = c.kh * c.kw;
    std::vector<float> scale(c.cout_pad, 0.f), bias(c.cout_pad, 0.f);
    for (int o = 0; o < c.cout; ++o) {
        float cb = b ? b->data[o] : 0.f;
        if (!bn.empty()) {
            const float* g = d.get(bn + ".weight")->data;
            const float* be = d.get(bn + ".bias")->data;
            const float* mu = d.get(bn + ".running_mean")->data;
            const float* var = d.get(bn + ".running_var")->data;
            float s = g[o] / sqrtf(var[o] + 1e-5f);
            scale[o] = s;
            bias[o] = be[o] + (cb - mu[o]) * s;
        } else {
            scale[o] = 1.f;
            bias[o] = cb;
        }
    }
    std::vector<float> wf((size_t)taps * c.cin * c.cout_pad, 0.f);
    std::vector<__nv_bfloat16> wb((size_t)taps * c.cout_pad * c.cin, __float2bfloat16(0.f));
    for (int o = 0; o < c.cout; ++o)
        for (int i = 0; i < c.cin; ++i)
            for (int t = 0; t < taps; ++t) {
                float v = w->data[((int64_t)o * c.cin + i) * taps + t];
                wf[((size_t)t * c.cin + i) * c.cout_pad + o] = v;
                wb[((size_t)t * c.cout_pad + o) * c.cin + i] = __float2bfloat16(v);
            }
    c.w_f32 = to_device(h, wf);
    c.w_bf16 = to_device(h, wb);
    c.scale = to_device(h, scale);
    c.bias = to_device(h, bias);
    if (split) {      // [tap][cout_pad][hi(cin) | lo(cin)] : w = bf16 hi + bf16 lo (conv_tc.cu, PAIR stages)
        std::vector<__nv_bfloat16> ws((size_t)taps * c.cout_pad * 2 * c.cin, __float2bfloat16(0.f));
        for (int o = 0; o < c.cout; ++o)
            for (int i = 0; i < c.cin; ++i)
                for (int t = 0; t < taps; ++t) {
                    float v = w->data[((int64_t)o * c.cin + i) * taps + t];
                    __nv_bfloat16 hi = __float2bfloat16(v), lo = __float2bfloat16(v - __bfloat162float(hi));
                    size_t base = ((size_t)t * c.cout_pad + o) * 2 * c.cin;
                    ws[base + i] = hi;
                    ws[base + c.cin + i] = lo;
                }
        c.w_split = to_device(h, ws);
    }
    return c;
}

// nn.Linear / LSTM input projection as a 1x1 convolution: rows of `mats` are stacked along cout.
ConvW make_linear(Handle* h, const std::vector<const bbocr_tensor*>& mats, const std::vector<std::vector<const bbocr_tensor*>>& biases,
                  bool split = true) {
    ConvW c;
    c.cin = (int)mats[0]->shape[1];
    c.cout = 0;
    for (auto* m : mats) {
        ARG_CHECK(m->ndim == 2 && m->shape[1] == c.cin, "linear: shape");
        c.cout += (int)m->shape[0];
    }
    c.kh = c.kw = 1; c.pad = 0; c.dil = 1;
    c.cout_pad = (c.cout + 15) / 16 * 16;
    std::vector<float> scale(c.cout_pad, 0.f), bias(c.cout_pad, 0.f);
    std::vector<float> wf((size_t)c.cin * c.cout_pad, 0.f);
    std::vector<__nv_bfloat16> wb((size_t)c.cout_pad * c.cin, __float2bfloat16(0.f));
    std::vector<__nv_bfloat16> ws(split ? (size_t)c.cout_pad * 2 * c.cin : 0, __float2bfloat16(0.f));
    int o0 = 0;
    for (size_t mi = 0; mi < mats.size(); ++mi) {
        const bbocr_tensor* m = mats[mi];
        int rows = (int)m->shape[0];
        for (int o = 0; o < rows; ++o) {
            scale[o0 + o] = 1.f;
            float bsum = 0.f;
            for (auto* bt : biases[mi]) bsum += bt->data[o];
            bias[o0 + o] = bsum;
            for (int i = 0; i < c.cin; ++i) {
                float v = m->data[(int64_t)o * c.cin + i];
                wf[(size_t)i * c.cout_pad + o0 + o] = v;
                wb[(size_t)(o0 + o) * c.cin + i] = __float2bfloat16(v);
                if (split) {
                    __nv_bfloat16 hi = __float2bfloat16(v), lo = __float2bfloat16(v - __bfloat162float(hi));
                    size_t base = (size_t)(o0 + o) * 2 * c.cin;
                    ws[base + i] = hi;
                    ws[base + c.cin + i] = lo;
                }
            }
        }
        o0 += rows;
    }
    c.w_f32 = to_device(h, wf);
    c.w_bf16 = to_device(h, wb);
    if (split) c.w_split = to_device(h, ws);
    c.scale = to_device(h, scale);
    c.bias = to_device(h, bias);
    return c;
}

LstmW make_lstm(Handle* h, const Dict& d, const std::string& p) {
    LstmW l;
    const bbocr_tensor* wih_f = d.get(p + "rnn.weight_ih_l0");
    const bbocr_tensor* wih_b = d.get(p + "rnn.weight_ih_l0_reverse");
    l.in_proj = make_linear(h, {wih_f, wih_b},
                            {{d.get(p + "rnn.bias_ih_l0"), d.get(p + "rnn.bias_hh_l0")},
                             {d.get(p + "rnn.bias_ih_l0_reverse"), d.get(p + "rnn.bias_hh_l0_reverse")}});
    ARG_CHECK(l.in_proj.cout == 2048 && l.in_proj.cin == 256, "LSTM must be 256->256 bidirectional");
    std::vector<float> whh((size_t)2 * 256 * 1024);
    const bbocr_tensor* hh[2] = {d.get(p + "rnn.weight_hh_l0"), d.get(p + "rnn.weight_hh_l0_reverse")};
    for (int dir = 0; dir < 2; ++dir) {
        ARG_CHECK(numel(hh[dir]) == 1024 * 256, "weight_hh shape");
        for (int r = 0; r < 1024; ++r)
            for (int k = 0; k < 256; ++k) whh[((size_t)dir * 256 + k) * 1024 + r] = hh[dir]->data[(int64_t)r * 256 + k];
    }
    l.w_hh = to_device(h, whh);
    l.linear = make_linear(h, {d.get(p + "linear.weight")}, {{d.get(p + "linear.bias")}});
    return l;
}

}  // namespace

// test hook: a bias-only convolution from raw PyTorch-layout arrays
ConvW make_conv_raw(Handle* h, const float* w, const float* bias, int cout, int cin, int kh, int kw, int pad, int dil) {
    bbocr_tensor t[2];
    t[0].name = "c.weight"; t[0].data = w; t[0].ndim = 4;
    t[0].shape[0] = cout; t[0].shape[1] = cin; t[0].shape[2] = kh; t[0].shape[3] = kw;
    t[1].name = "c.bias"; t[1].data = bias; t[1].ndim = 1; t[1].shape[0] = cout;
    Dict d(t, bias ? 2 : 1);
    return make_conv(h, d, "c", "", pad, dil, true);
}

void load_craft(Handle* h, const bbocr_tensor* t, int n) {
    Dict d(t, n);
    CraftW& c = h->craft;
    auto vgg = [&](int slice_conv, int idx, int slice_bn) {
        return make_conv(h, d, "basenet.slice" + std::to_string(slice_conv) + "." + std::to_string(idx),
                         "basenet.slice" + std::to_string(slice_bn) + "." + std::to_string(idx + 1), 1, 1, true);
    };
    c.c1_1 = vgg(1, 0, 1);  c.c1_2 = vgg(1, 3, 1);  c.c2_1 = vgg(1, 7, 1);  c.c2_2 = vgg(1, 10, 1);
    c.c3_1 = vgg(2, 14, 2); c.c3_2 = vgg(2, 17, 2);
    c.c3_3 = vgg(3, 20, 3); c.c4_1 = vgg(3, 24, 3); c.c4_2 = vgg(3, 27, 3);
    c.c4_3 = vgg(4, 30, 4); c.c5_1 = vgg(4, 34, 4); c.c5_2 = vgg(4, 37, 4);
    c.fc6 = make_conv(h, d, "basenet.slice5.1", "", 6, 6, true);
    c.fc7 = make_conv(h, d, "basenet.slice5.2", "", 0, 1, true);
    ConvW* ups[4][2] = {{&c.up1a, &c.up1b}, {&c.up2a, &c.up2b}, {&c.up3a, &c.up3b}, {&c.up4a, &c.up4b}};
    for (int i = 0; i < 4; ++i) {
        std::string p = "upconv" + std::to_string(i + 1) + ".conv.";
        *ups[i][0] = make_conv(h, d, p + "0", p + "1", 0, 1, true);
        *ups[i][1] = make_conv(h, d, p + "3", p + "4", 1, 1, true);
    }
    c.cls0 = make_conv(h, d, "conv_cls.0", "", 1, 1, true);
    c.cls1 = make_conv(h, d, "conv_cls.2", "", 1, 1, true);
    c.cls2 = make_conv(h, d, "conv_cls.4", "", 1, 1, true);
    c.cls3 = make_conv(h, d, "conv_cls.6", "", 0, 1);
    c.cls4 = make_conv(h, d, "conv_cls.8", "", 0, 1);
    ARG_CHECK(c.c1_1.cin == 3 && c.c1_1.cout == 64 && c.cls4.cout == 2, "CRAFT: unexpected shapes");
    {
        // conv1_1 for the tensor-core path: the 3x3x3 neighbourhood is gathered into 32 channels (27 + 5 zeros) by
        // k_im2col_rgb, which turns the layer into a 1x1 convolution with Cin = 32: W32[o][tap*3 + c] = w[o][c][tap]
        const bbocr_tensor* w = d.get("basenet.slice1.0.weight");
        std::vector<__nv_bfloat16> wb((size_t)64 * 32, __float2bfloat16(0.f)), ws((size_t)64 * 64, __float2bfloat16(0.f));
        std::vector<float> wf((size_t)32 * 64, 0.f);
        for (int o = 0; o < 64; ++o)
            for (int ci = 0; ci < 3; ++ci)
                for (int t = 0; t < 9; ++t) {
                    float v = w->data[((int64_t)o * 3 + ci) * 9 + t];
                    __nv_bfloat16 hi = __float2bfloat16(v);
                    wb[(size_t)o * 32 + t * 3 + ci] = hi;
                    ws[(size_t)o * 64 + t * 3 + ci] = hi;
                    ws[(size_t)o * 64 + 32 + t * 3 + ci] = __float2bfloat16(v - __bfloat162float(hi));
                    wf[(size_t)(t * 3 + ci) * 64 + o] = v;
                }
        ConvW e = c.c1_1;
        e.cin = 32; e.kh = e.kw = 1; e.pad = 0; e.dil = 1;
        e.w_bf16 = to_device(h, wb);
        e.w_split = to_device(h, ws);
        e.w_f32 = to_device(h, wf);
        c.c1_1_tc = e;
    }
    CUDA_CHECK(cudaDeviceSynchronize());      // pageable cudaMemcpy may return before its DMA lands; the lanes' streams are non-blocking
    h->craft_loaded = true;
}

void load_crnn(Handle* h, const bbocr_tensor* t, int n) {
    Dict d(t, n);
    CrnnW& c = h->crnn;
    const std::string p = "FeatureExtraction.ConvNet.";
    c.c0 = make_conv(h, d, p + "0", "", 1, 1);
    c.c1 = make_conv(h, d, p + "3", "", 1, 1, true);
    c.c2 = make_conv(h, d, p + "6", "", 1, 1, true);
    c.c3 = make_conv(h, d, p + "8", "", 1, 1, true);
    c.c4 = make_conv(h, d, p + "11", p + "12", 1, 1, true);
    c.c5 = make_conv(h, d, p + "14", p + "15", 1, 1, true);
    c.c6 = make_conv(h, d, p + "18", "", 0, 1, true);
    c.l0 = make_lstm(h, d, "SequenceModeling.0.");
    c.l1 = make_lstm(h, d, "SequenceModeling.1.");
    c.pred = make_linear(h, {d.get("Prediction.weight")}, {{d.get("Prediction.bias")}});
    c.num_class = c.pred.cout;
    ARG_CHECK(c.c0.cin == 1 && c.c0.cout == 32 && c.c6.kh == 2 && c.pred.cin == 256, "CRNN: unexpected shapes");
    CUDA_CHECK(cudaDeviceSynchronize());
    h->crnn_loaded = true;
}

}  // namespace bbocr
