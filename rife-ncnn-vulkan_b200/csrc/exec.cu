// exec.cu -- see exec.h.  Shape rules and operator semantics cite the reference's generic layers
// (/root/reference/src/ncnn/src/layer/*.cpp); scheduling, memory planning and fusion are ours.
#include "exec.h"

#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "kernels.h"
#include "tc_conv.h"

namespace rife {

#define CUDA_OK(x)                                                                      \
    do {                                                                                \
        cudaError_t e_ = (x);                                                           \
        if (e_ != cudaSuccess) {                                                        \
            err = std::string(#x) + ": " + cudaGetErrorString(e_);                      \
            return -10;                                                                 \
        }                                                                               \
    } while (0)

static float* upload(const std::vector<float>& v, std::string& err) {
    if (v.empty()) return nullptr;
    float* d = nullptr;
    if (cudaMalloc(&d, v.size() * 4) != cudaSuccess) { err = "cudaMalloc failed"; return nullptr; }
    cudaMemcpy(d, v.data(), v.size() * 4, cudaMemcpyHostToDevice);
    return d;
}

NetRunner::~NetRunner() {
    if (owns_) for (auto& w : dw_own_) {
        cudaFree(w.wT);
        cudaFree(w.bias);
        cudaFree(w.slope);
        cudaFree(w.wpk);
        cudaFree(w.biasN);
    }
    for (auto& kv : plans_) cudaFree(kv.second->arena);
    cudaFree(pool_scratch_);
}

void NetRunner::clear_plans() {
    for (auto& kv : plans_) cudaFree(kv.second->arena);
    plans_.clear();
}

size_t NetRunner::arena_bytes() const {
    size_t s = 0;
    for (auto& kv : plans_) s += kv.second->arena_size;
    return s;
}

int NetRunner::init(const Net* net, std::string& err) {
    net_ = net;
    dw_own_.assign(net->layers.size(), DeviceWeights());
    dwp_ = &dw_own_;
    owns_ = true;
    for (size_t li = 0; li < net->layers.size(); li++) {
        const Layer& L = net->layers[li];
        DeviceWeights& W = dw_own_[li];
        if (L.type == "Convolution") {
            int cout = L.geti(0, 0), kw = L.geti(1, 0), kh = L.geti(11, kw);
            int kk = kw * kh;
            if (cout <= 0 || kk <= 0 || L.weight.size() % ((size_t)cout * kk)) { err = "bad conv weights in " + L.name; return -4; }
            int cin = (int)(L.weight.size() / ((size_t)cout * kk));
            int ocpad = (cout + 63) / 64 * 64;
            std::vector<float> t((size_t)cin * kk * ocpad, 0.f);
            for (int oc = 0; oc < cout; oc++)      // convolution.cpp:179: weights [oc][ic][kk]
                for (int ic = 0; ic < cin; ic++)
                    for (int k = 0; k < kk; k++) t[((size_t)ic * kk + k) * ocpad + oc] = L.weight[((size_t)oc * cin + ic) * kk + k];
            W.wT = upload(t, err);
            W.ocpad = ocpad;
        } else if (L.type == "Deconvolution") {
            int cout = L.geti(0, 0), kw = L.geti(1, 0), kh = L.geti(11, kw);
            if (kw != 4 || kh != 4 || L.geti(3, 1) != 2 || L.geti(4, 0) != 1) { err = "unsupported deconvolution shape in " + L.name; return -4; }
            if (cout <= 0 || L.weight.empty() || L.weight.size() % ((size_t)cout * 16)) { err = "bad deconv weights in " + L.name; return -4; }
            int cin = (int)(L.weight.size() / ((size_t)cout * 16));
            int ocpad = (cout + 63) / 64 * 64;
            // deconvolution.cpp:68-141: out[y = i*2 + ky - 1] += in[i] * w[oc][ic][ky][kx].  Gather form per output
            // parity (py,px): taps r,c in {0,1} read input (a - 1 + py + r, b - 1 + px + c) with ky = 3 - py - 2r.
            std::vector<float> t((size_t)4 * cin * 4 * ocpad, 0.f);
            for (int p = 0; p < 4; p++) {
                int py = p >> 1, px = p & 1;
                for (int ic = 0; ic < cin; ic++)
                    for (int r = 0; r < 2; r++)
                        for (int c = 0; c < 2; c++) {
                            int ky = 3 - py - 2 * r, kx = 3 - px - 2 * c;
                            for (int oc = 0; oc < cout; oc++)
                                t[(((size_t)p * cin + ic) * 4 + r * 2 + c) * ocpad + oc] = L.weight[((size_t)oc * cin + ic) * 16 + ky * 4 + kx];
                        }
            }
            W.wT = upload(t, err);
            W.ocpad = ocpad;
        } else if (L.type == "InnerProduct") {
            W.wT = upload(L.weight, err);
        }
        // tensor-core eligibility (tc_conv.cu): 3x3 s1 / s2 p1 conv with Cin % 16 == 0 (narrow stride-2 inputs are padded to 16) and
        // N in {32,48,64,96,128,192}, wider ones (256 / 384 / 512) as slices of 128; 5x5 s1 p2 with N in {48,96,128,192};
        // deconv 4x4 s2 p1 re-expressed as a 3x3 conv with N = 4 * ocs in {32, 96}, wider ones as slices of 16 channels
        if ((L.type == "Convolution" || L.type == "Deconvolution") && L.weight_is_fp16) {
            int cout = L.geti(0, 0), k = L.geti(1, 0);
            bool isconv = L.type == "Convolution";
            const bool is5 = isconv && k == 5 && L.geti(11, 5) == 5;
            int kk = isconv ? (is5 ? 25 : 9) : 16;
            int cin = (int)(L.weight.size() / ((size_t)cout * kk));
            int N = 0, ocs = 0, s2 = 0, cinp = cin, k5 = 0;
            // 5x5 s1 p2 (the residual blocks of the rife / HD / UHD / anime flownets): one kernel row per pipeline stage
            if (is5 && L.geti(3, 1) == 1 && L.geti(4, 0) == 2 && L.geti(2, 1) == 1 && (cout == 48 || cout == 96 || cout == 128 || cout == 192)) { N = cout; k5 = 1; }
            // (8-channel heads: padded to 16 GEMM columns; the C8 storage of a blob is allocated in multiples of 16 channels)
            if (isconv && k == 3 && L.geti(3, 1) == 1 && L.geti(4, 0) == 1 && L.geti(2, 1) == 1) N = cout == 8 ? 16 : cout;
            if (isconv && k == 3 && L.geti(3, 1) == 2 && L.geti(4, 0) == 1 && L.geti(2, 1) == 1) { N = cout; s2 = 1; if (cin < 16) cinp = 16; }
            if (!isconv) { ocs = (cout + 7) / 8 * 8; N = 4 * ocs; }
            // Layers wider than one accumulator set run as output-channel slices: convolutions with 256 / 384 / 512 output
            // channels as slices of 128 (the context / fusion pyramids of the 3-net families), deconvolutions with 32 .. 256
            // output channels as slices of 16 (4 parities x 16 = 64 GEMM columns; the deconv epilogue holds a row of them in registers)
            int nchunks = 1, cchunk = cout;  // output channels per slice (the last slice may be narrower: zero-padded columns)
            if (isconv && !k5 && N > 192 && N % 128 == 0 && N <= 512) { nchunks = N / 128; cchunk = 128; N = 128; }
            if (!isconv && !(N == 32 || N == 96) && cout % 8 == 0 && cout >= 32 && cout <= 256) { cchunk = 24; nchunks = (cout + 23) / 24; ocs = 24; N = 96; }
            bool nok = isconv ? (N == 16 || N == 32 || N == 48 || N == 64 || N == 96 || N == 128 || N == 192) : (N == 32 || N == 96);
            if (nok && cinp % 16 == 0 && cinp >= 16 && (size_t)cin * cout * kk == L.weight.size()) {
                std::vector<uint16_t> pk_all;
                size_t chunk_elems = 0;
                for (int ch = 0; ch < nchunks; ch++) {
                    std::vector<uint16_t> pk;
                    const float* wc = L.weight.data() + (size_t)ch * cchunk * cin * kk;  // weights are [oc][ic][kk]: a slice of output channels is contiguous
                    const int cn = std::min(cchunk, cout - ch * cchunk);                // channels of this slice
                    if (k5) pack_conv5x5_weights(wc, cn, cin, N, pk);
                    else if (isconv && s2) pack_conv3x3s2_weights(wc, cn, cin, cinp, N, pk);
                    else if (isconv) pack_conv3x3_weights(wc, cn, cin, N, pk);
                    else pack_deconv4x4_weights(wc, cn, cin, ocs, N, pk);
                    chunk_elems = pk.size();
                    pk_all.insert(pk_all.end(), pk.begin(), pk.end());
                }
                std::vector<float> bN((size_t)N * nchunks, 0.f);
                if (!L.bias.empty()) {
                    for (int ch = 0; ch < nchunks; ch++) {
                        const int cn = std::min(cchunk, cout - ch * cchunk);
                        if (isconv) for (int i = 0; i < cn; i++) bN[(size_t)ch * N + i] = L.bias[ch * cchunk + i];
                        else for (int p = 0; p < 4; p++) for (int i = 0; i < cn; i++) bN[(size_t)ch * N + p * ocs + i] = L.bias[ch * cchunk + i];
                    }
                }
                if (cudaMalloc(&W.wpk, pk_all.size() * 2) != cudaSuccess) { err = "cudaMalloc failed"; return -5; }
                cudaMemcpy(W.wpk, pk_all.data(), pk_all.size() * 2, cudaMemcpyHostToDevice);
                W.biasN = upload(bN, err);
                W.tcN = N; W.ocs = ocs; W.cin = cin; W.cinp = cinp; W.tc_s2 = s2; W.tc_k5 = k5; W.nchunks = nchunks; W.chunk_elems = chunk_elems;
            }
        }
        if (!L.bias.empty()) W.bias = upload(L.bias, err);
        if (!L.slope.empty()) W.slope = upload(L.slope, err);
        if (!err.empty()) return -5;
    }
    return 0;
}

namespace {
struct FreeList {
    std::vector<std::pair<size_t, size_t>> free;  // (offset, size), sorted by offset
    size_t top = 0;
    size_t alloc(size_t n) {
        n = (n + 255) & ~(size_t)255;
        for (size_t i = 0; i < free.size(); i++)
            if (free[i].second >= n) {
                size_t off = free[i].first;
                if (free[i].second == n) free.erase(free.begin() + i);
                else { free[i].first += n; free[i].second -= n; }
                return off;
            }
        // grow: extend a trailing free block if it touches the top
        if (!free.empty() && free.back().first + free.back().second == top) {
            size_t off = free.back().first;
            top = off + n;
            free.pop_back();
            return off;
        }
        size_t off = top;
        top += n;
        return off;
    }
    void release(size_t off, size_t n) {
        n = (n + 255) & ~(size_t)255;
        auto it = std::lower_bound(free.begin(), free.end(), std::make_pair(off, (size_t)0));
        it = free.insert(it, {off, n});
        size_t i = it - free.begin();
        if (i + 1 < free.size() && free[i].first + free[i].second == free[i + 1].first) {
            free[i].second += free[i + 1].second;
            free.erase(free.begin() + i + 1);
        }
        if (i > 0 && free[i - 1].first + free[i - 1].second == free[i].first) {
            free[i - 1].second += free[i].second;
            free.erase(free.begin() + i);
        }
    }
};
}  // namespace

int NetRunner::build_plan(const std::vector<std::pair<std::string, Tensor>>& inputs, const std::vector<std::string>& outputs, Plan& plan, std::string& err) {
    const Net& net = *net_;
    const int nb = (int)net.blob_names.size(), nl = (int)net.layers.size();
    plan.blobs.assign(nb, Tensor());
    plan.external_slot.assign(nb, -1);
    for (size_t i = 0; i < inputs.size(); i++) {
        int b = net.find_blob(inputs[i].first);
        if (b < 0) { err = "unknown input blob " + inputs[i].first; return -20; }
        plan.external_slot[b] = (int)i;
        plan.blobs[b] = inputs[i].second;
    }
    // 1. needed layers
    std::vector<char> need(nl, 0);
    std::vector<int> stack;
    for (auto& o : outputs) {
        int b = net.find_blob(o);
        if (b < 0) { err = "unknown output blob " + o; return -20; }
        plan.out_ids.push_back(b);
        stack.push_back(b);
    }
    while (!stack.empty()) {
        int b = stack.back();
        stack.pop_back();
        if (plan.external_slot[b] >= 0) continue;
        int l = net.producer[b];
        if (l < 0) { err = "blob without producer: " + net.blob_names[b]; return -20; }
        if (need[l]) continue;
        need[l] = 1;
        if (net.layers[l].type == "Input") { err = "missing input blob " + net.blob_names[b]; return -21; }
        for (int bb : net.layers[l].bottoms) stack.push_back(bb);
    }
    // 1b. arity: everything below indexes bottoms / tops by position; a damaged .param can name fewer (or more) than the type takes
    for (int l = 0; l < nl; l++) {
        if (!need[l]) continue;
        const Layer& L = net.layers[l];
        size_t lo = 1, hi = 1;  // bottoms
        if (L.type == "Concat") hi = (size_t)-1;
        else if (L.type == "BinaryOp") hi = 2;
        else if (L.type == "Eltwise" || L.type == "rife.Warp") lo = hi = 2;
        const bool tops_ok = L.type == "Split" ? L.tops.size() >= 1 : L.tops.size() == 1;
        if (L.bottoms.size() < lo || L.bottoms.size() > hi || !tops_ok) { err = "wrong number of inputs / outputs for " + L.type + " layer " + L.name; return -23; }
        if (L.type == "PReLU" && L.slope.empty()) { err = "PReLU without slope data: " + L.name; return -23; }
    }
    // 2. consumer counts among needed layers (+1 for requested outputs)
    std::vector<int> ncons(nb, 0);
    for (int l = 0; l < nl; l++)
        if (need[l])
            for (int b : net.layers[l].bottoms) ncons[b]++;
    for (int b : plan.out_ids) ncons[b] += 1000;
    auto sole_consumer = [&](int b) -> int {
        if (ncons[b] != 1) return -1;
        for (int l = 0; l < nl; l++)
            if (need[l])
                for (int bb : net.layers[l].bottoms)
                    if (bb == b) return l;
        return -1;
    };
    // 3. steps with epilogue fusion
    std::vector<char> skipped(nl, 0);
    for (int l = 0; l < nl; l++) {
        if (!need[l] || skipped[l]) continue;
        const Layer& L = net.layers[l];
        Step s;
        s.layer = l;
        s.out_blob = L.tops.empty() ? -1 : L.tops[0];
        if (fuse && (L.type == "Convolution" || L.type == "Deconvolution") && plan.external_slot[L.tops[0]] < 0) {
            int t = L.tops[0];
            int c1 = sole_consumer(t);
            if (c1 >= 0 && plan.external_slot[net.layers[c1].tops[0]] < 0) {
                const Layer& C1 = net.layers[c1];
                if (C1.type == "PReLU" && (C1.slope.size() == 1 || C1.slope.size() == (size_t)std::max(0, L.geti(0, 0)))) {  // (any other slope count is refused by the shape pass)
                    s.fused_act_layer = c1;
                    s.out_blob = C1.tops[0];
                    skipped[c1] = 1;
                } else if (L.type == "Convolution" && C1.type == "BinaryOp" && C1.bottoms.size() == 2 && C1.geti(0, 0) == 0 && C1.geti(1, 0) == 0 &&
                           C1.bottoms[0] != C1.bottoms[1]) {
                    int u = C1.tops[0];
                    int c2 = sole_consumer(u);
                    if (c2 >= 0 && net.layers[c2].type == "ReLU" && plan.external_slot[net.layers[c2].tops[0]] < 0) {
                        s.fused_add_blob = C1.bottoms[0] == t ? C1.bottoms[1] : C1.bottoms[0];
                        s.fused_act_layer = c2;
                        s.out_blob = net.layers[c2].tops[0];
                        skipped[c1] = skipped[c2] = 1;
                    }
                }
            }
        }
        if (tc_mode > 0 && (*dwp_)[l].wpk && plan.external_slot[L.tops[0]] < 0) {
            int own_act = L.geti(9, 0);
            if (L.type == "Convolution") {
                bool ok = own_act == 0 || (own_act == 2 && s.fused_act_layer < 0 && s.fused_add_blob < 0);
                if (ok) s.kind = 1;
            } else if ((own_act == 0 || own_act == 4) && (s.fused_act_layer < 0 || own_act == 0)) {
                s.kind = 1;  // (a PReLU / leaky fused behind the deconvolution is applied by its epilogue)
                int c1 = sole_consumer(L.tops[0]);
                if (c1 >= 0 && net.layers[c1].type == "PixelShuffle" && net.layers[c1].geti(0, 1) == 2 && net.layers[c1].geti(1, 0) == 0 &&
                    L.geti(0, 0) % 4 == 0 && plan.external_slot[net.layers[c1].tops[0]] < 0 && (*dwp_)[l].nchunks == 1 && s.fused_act_layer < 0) {
                    s.fused_ps_layer = c1;
                    s.out_blob = net.layers[c1].tops[0];
                    skipped[c1] = 1;
                }
            }
        }
        plan.steps.push_back(s);
    }
    // 4. shape inference + alias analysis
    std::vector<int>& root = plan.root;
    std::vector<size_t>& eoff = plan.eoff;
    root.assign(nb, -1);
    eoff.assign(nb, 0);
    for (int b = 0; b < nb; b++)
        if (plan.external_slot[b] >= 0) root[b] = b;
    auto same_shape = [](const Tensor& a, const Tensor& b) { return a.dims == b.dims && a.c == b.c && a.h == b.h && a.w == b.w; };
    for (const Step& s : plan.steps) {
        const Layer& L = net.layers[s.layer];
        auto in = [&](int i) -> const Tensor& { return plan.blobs[L.bottoms[i]]; };
        for (int b : L.bottoms)
            if (plan.blobs[b].dims == 0) { err = "internal: blob " + net.blob_names[b] + " used before defined in " + L.name; return -22; }
        Tensor o;
        o.dims = 3;
        const std::string& T = L.type;
        if (T == "Split") {
            for (int t : L.tops) { plan.blobs[t] = in(0); root[t] = root[L.bottoms[0]]; eoff[t] = eoff[L.bottoms[0]]; }
            continue;
        } else if (T == "Crop") {
            const ParamVal* st = L.get(9);
            const ParamVal* en = L.get(10);
            const ParamVal* ax = L.get(11);
            if (!st || !en || !ax || ax->ai.size() != 1 || ax->ai[0] != 0 || st->ai.size() != 1 || en->ai.size() != 1 || in(0).dims != 3) { err = "unsupported Crop form in " + L.name; return -23; }
            int s0 = st->ai[0], e0 = std::min(en->ai[0], in(0).c);  // crop.cpp:387-430 (end clamps to the extent)
            if (s0 < 0 || s0 >= e0) { err = "bad Crop range in " + L.name; return -23; }
            o = in(0);
            o.c = e0 - s0;
            plan.blobs[L.tops[0]] = o;
            root[L.tops[0]] = root[L.bottoms[0]];
            eoff[L.tops[0]] = eoff[L.bottoms[0]] + (size_t)s0 * in(0).h * in(0).w;
            continue;
        } else if (T == "Concat") {
            if (L.geti(0, 0) != 0) { err = "Concat axis != 0 in " + L.name; return -23; }
            o = in(0);
            o.c = 0;
            for (size_t i = 0; i < L.bottoms.size(); i++) {
                if (in(i).h != in(0).h || in(i).w != in(0).w) { err = "Concat shape mismatch in " + L.name; return -23; }
                o.c += in(i).c;
            }
        } else if (T == "Convolution") {
            int k = L.geti(1, 0), sd = L.geti(3, 1), pad = L.geti(4, 0);
            if (L.geti(11, k) != k || L.geti(13, sd) != sd || L.geti(2, 1) != 1 || k <= 0 || sd <= 0) { err = "unsupported conv form in " + L.name; return -23; }
            o.c = L.geti(0, 0);
            // the kernels index the weights by the input's channel count: it must be the one the weights were stored for
            if (o.c <= 0 || (size_t)in(0).c * o.c * k * k != L.weight.size() || in(0).dims != 3) { err = "conv input channels do not match the weights in " + L.name; return -23; }
            o.h = (in(0).h + 2 * pad - k) / sd + 1;  // convolution.cpp: outh = (h - kernel_extent) / stride + 1 after padding
            o.w = (in(0).w + 2 * pad - k) / sd + 1;
        } else if (T == "Deconvolution") {
            o.c = L.geti(0, 0);
            if (o.c <= 0 || (size_t)in(0).c * o.c * 16 != L.weight.size() || in(0).dims != 3) { err = "deconv input channels do not match the weights in " + L.name; return -23; }
            o.h = (in(0).h - 1) * 2 + 4 - 2;  // deconvolution.cpp:183-188 (crop by pad on every side)
            o.w = (in(0).w - 1) * 2 + 4 - 2;
        } else if (T == "Interp") {
            if (L.geti(0, 0) != 2 || L.geti(3, 0) != 0 || L.geti(4, 0) != 0 || L.geti(6, 0) != 0) { err = "unsupported Interp form in " + L.name; return -23; }
            o = in(0);
            o.h = (int)(in(0).h * L.getf(1, 1.f));  // interp.cpp:438-441
            o.w = (int)(in(0).w * L.getf(2, 1.f));
        } else if (T == "PixelShuffle") {
            int r = L.geti(0, 1);
            if (L.geti(1, 0) != 0) { err = "PixelShuffle mode 1 unsupported"; return -23; }
            o.c = in(0).c / (r * r);
            o.h = in(0).h * r;
            o.w = in(0).w * r;
        } else if (T == "BinaryOp") {
            if (L.bottoms.size() == 2) {
                const Tensor& a = in(0);
                const Tensor& b = in(1);
                o = a.count() >= b.count() ? a : b;
            } else o = in(0);
        } else if (T == "Eltwise") {
            if (L.bottoms.size() != 2 || L.geti(0, 0) != 1 || !same_shape(in(0), in(1))) { err = "unsupported Eltwise form in " + L.name; return -23; }
            o = in(0);
        } else if (T == "ReLU" || T == "PReLU" || T == "Sigmoid" || T == "Clip" || T == "UnaryOp") {
            if (T == "PReLU" && L.slope.size() != 1 && (in(0).dims != 3 || L.slope.size() != (size_t)in(0).c)) { err = "PReLU slope count does not match the input channels in " + L.name; return -23; }
            o = in(0);
        } else if (T == "rife.Warp") {
            if (in(1).c < 2 || in(1).h != in(0).h || in(1).w != in(0).w) { err = "Warp shape mismatch in " + L.name; return -23; }
            o = in(0);
        } else if (T == "Pooling") {
            if (L.geti(0, 0) != 1 || L.geti(4, 0) != 1) { err = "unsupported Pooling form in " + L.name; return -23; }
            o.dims = 1;
            o.w = in(0).c;
            o.c = o.h = 1;
        } else if (T == "InnerProduct") {
            if (L.geti(0, 0) <= 0 || in(0).count() * (size_t)L.geti(0, 0) != L.weight.size()) { err = "InnerProduct input size does not match the weights in " + L.name; return -23; }
            o.dims = 1;
            o.w = L.geti(0, 0);
            o.c = o.h = 1;
        } else {
            err = "unsupported layer type " + T + " (" + L.name + ")";
            return -24;
        }
        if (s.fused_ps_layer >= 0) { o.c /= 4; o.h *= 2; o.w *= 2; }
        int ob = s.out_blob;
        plan.blobs[ob] = o;
        root[ob] = ob;
        eoff[ob] = 0;
        if (s.fused_add_blob >= 0 && !same_shape(plan.blobs[s.fused_add_blob], o)) { err = "internal: fused residual shape mismatch at " + L.name; return -22; }
    }
    // 5. storage formats: a root blob may exist as planar fp32 and/or C8 fp16 (tensor-core layout); insert the
    //    conversions a consumer needs, then do liveness + arena assignment over (root, format) storages
    plan.split = tc_mode == 1;
    {
        // a tensor-core step needs whole-blob inputs (no channel-offset aliases); a stride-2 step needs even input
        // dims and is the only tensor-core consumer format of its input (space-to-depth C8)
        plan.c8_s2d.assign(nb, 0);
        std::vector<char> c8_used_s1(nb, 0);
        for (Step& s : plan.steps) {
            if (s.kind != 1) continue;
            const Layer& L = net.layers[s.layer];
            const DeviceWeights& W = (*dwp_)[s.layer];
            const Tensor& x = plan.blobs[L.bottoms[0]];
            bool ok = eoff[L.bottoms[0]] == 0 && x.dims == 3 && x.c == W.cin;
            if (s.fused_add_blob >= 0 && eoff[s.fused_add_blob] != 0) ok = false;
            if (s.fused_add_blob >= 0 && plan.blobs[root[s.fused_add_blob]].c != plan.blobs[s.fused_add_blob].c) ok = false;
            if (plan.blobs[root[L.bottoms[0]]].c != x.c) ok = false;
            if (W.tc_s2 && ((x.h | x.w) & 1)) ok = false;
            if (W.cinp != W.cin && tc_mode != 1) ok = false;  // narrow block-head inputs carry flow: split precision only
            int r = root[L.bottoms[0]];
            if (ok && W.tc_s2 && c8_used_s1[r]) ok = false;
            if (ok && !W.tc_s2 && plan.c8_s2d[r]) ok = false;
            if (!ok) {
                if (s.fused_ps_layer >= 0) { err = "internal: cannot undo pixelshuffle fusion at " + L.name; return -22; }
                s.kind = 0;
                continue;
            }
            if (W.tc_s2) plan.c8_s2d[r] = 1; else c8_used_s1[r] = 1;
            if (s.fused_add_blob >= 0) c8_used_s1[root[s.fused_add_blob]] = 1;
        }
        // a residual must be in plain C8 form: demote stride-2 consumers whose input doubles as a residual
        for (Step& s : plan.steps)
            if (s.kind == 1 && (*dwp_)[s.layer].tc_s2 && c8_used_s1[root[net.layers[s.layer].bottoms[0]]]) {
                plan.c8_s2d[root[net.layers[s.layer].bottoms[0]]] = 0;
                s.kind = 0;
            }
        std::vector<Step> final_steps;
        std::vector<char> have_planar(nb, 0), have_c8(nb, 0);
        for (int b = 0; b < nb; b++)
            if (plan.external_slot[b] >= 0) have_planar[b] = 1;
        auto require = [&](int blob, bool c8) {
            int r = root[blob];
            if (c8 && !have_c8[r]) {
                Step c; c.kind = 2; c.conv_root = r;
                final_steps.push_back(c);
                have_c8[r] = 1;
            } else if (!c8 && !have_planar[r]) {
                Step c; c.kind = 3; c.conv_root = r;
                final_steps.push_back(c);
                have_planar[r] = 1;
            }
        };
        for (const Step& s : plan.steps) {
            const Layer& L = net.layers[s.layer];
            bool tc = s.kind == 1;
            if (L.type == "Split" || L.type == "Crop") { final_steps.push_back(s); continue; }  // aliases: no data touched
            for (int b : L.bottoms) require(b, tc);
            if (s.fused_add_blob >= 0) require(s.fused_add_blob, tc);
            final_steps.push_back(s);
            if (L.type == "Split" || L.type == "Crop") continue;
            if (tc && L.type == "Convolution") have_c8[s.out_blob] = 1;
            else have_planar[s.out_blob] = 1;
        }
        for (int b : plan.out_ids) require(b, false);
        plan.steps.swap(final_steps);
    }
    const int ns = (int)plan.steps.size();
    // storage id = root * 2 + (c8 ? 1 : 0)
    std::vector<int> birth(2 * nb, -1), death(2 * nb, -1);
    auto touch = [&](int blob, bool c8, int i, bool write) {
        int sid = root[blob] * 2 + (c8 ? 1 : 0);
        if (write && birth[sid] < 0) birth[sid] = i;
        death[sid] = std::max(death[sid], i);
    };
    for (int i = 0; i < ns; i++) {
        const Step& s = plan.steps[i];
        if (s.kind == 2) { touch(s.conv_root, false, i, false); touch(s.conv_root, true, i, true); continue; }
        if (s.kind == 3) { touch(s.conv_root, true, i, false); touch(s.conv_root, false, i, true); continue; }
        const Layer& L = net.layers[s.layer];
        bool tc = s.kind == 1;
        if (L.type == "Split" || L.type == "Crop") continue;
        for (int b : L.bottoms) touch(b, tc, i, false);
        if (s.fused_add_blob >= 0) touch(s.fused_add_blob, tc, i, false);
        touch(s.out_blob, tc && L.type == "Convolution", i, true);
    }
    for (int b : plan.out_ids) death[root[b] * 2] = 1 << 30;
    plan.offset.assign(nb, (size_t)-1);
    plan.offset_c8.assign(nb, (size_t)-1);
    auto storage_bytes = [&](int sid) -> size_t {
        const Tensor& t = plan.blobs[sid / 2];
        const size_t cpad = (size_t)((t.c + 15) / 16 * 16);
        return (sid & 1) ? cpad * t.h * t.w * sizeof(uint16_t) * (plan.split ? 2 : 1) : t.count() * sizeof(float);
    };
    FreeList fl;
    std::vector<std::vector<int>> born_at(ns), free_at(ns);
    for (int sid = 0; sid < 2 * nb; sid++) {
        if (birth[sid] < 0) continue;  // external or never materialised
        born_at[birth[sid]].push_back(sid);
        if (death[sid] < ns) free_at[std::max(death[sid], birth[sid])].push_back(sid);
    }
    for (int i = 0; i < ns; i++) {
        for (int sid : born_at[i]) {
            size_t off = fl.alloc(storage_bytes(sid));
            if (sid & 1) plan.offset_c8[sid / 2] = off;
            else plan.offset[sid / 2] = off;
        }
        for (int sid : free_at[i]) fl.release((sid & 1) ? plan.offset_c8[sid / 2] : plan.offset[sid / 2], storage_bytes(sid));
    }
    plan.arena_size = fl.top;
    if (plan.arena_size) CUDA_OK(cudaMalloc(&plan.arena, plan.arena_size));
    return 0;
}

int NetRunner::run(const std::vector<std::pair<std::string, Tensor>>& inputs, const std::vector<std::string>& outputs, std::vector<Tensor>& out_tensors,
                   cudaStream_t st, std::string& err) {
    std::string key;
    char buf[96];
    for (auto& in : inputs) {
        snprintf(buf, sizeof buf, "%s:%d.%d.%d.%d;", in.first.c_str(), in.second.dims, in.second.c, in.second.h, in.second.w);
        key += buf;
    }
    key += "->";
    for (auto& o : outputs) key += o + ";";
    auto it = plans_.find(key);
    if (it == plans_.end()) {
        std::unique_ptr<Plan> p(new Plan());
        int r = build_plan(inputs, outputs, *p, err);
        if (r) return r;
        it = plans_.emplace(key, std::move(p)).first;
    }
    Plan& plan = *it->second;
    const std::vector<int>& root = plan.root;
    const std::vector<size_t>& eoff = plan.eoff;
    for (size_t b = 0; b < plan.blobs.size(); b++) {
        int r = root[b];
        if (r < 0) continue;
        float* base = plan.external_slot[r] >= 0 ? inputs[plan.external_slot[r]].second.p
                      : (plan.offset[r] != (size_t)-1 ? (float*)((char*)plan.arena + plan.offset[r]) : nullptr);
        plan.blobs[b].p = base ? base + eoff[b] : nullptr;
    }
    for (const Step& s : plan.steps) {
        int r = exec_step(plan, s, st, err);
        if (r) return r;
    }
    out_tensors.clear();
    for (int b : plan.out_ids) out_tensors.push_back(plan.blobs[b]);
    return 0;
}

int NetRunner::exec_step(Plan& plan, const Step& s, cudaStream_t st, std::string& err) {
    const Net& net = *net_;
    auto c8ptr = [&](int blob) -> __half* { return (__half*)((char*)plan.arena + plan.offset_c8[plan.root[blob]]); };
    if (s.kind == 2 || s.kind == 3) {
        const Tensor& t = plan.blobs[s.conv_root];
        const int cpad = (t.c + 15) / 16 * 16;
        if (s.kind == 2) launch_planar_to_c8(t.p, c8ptr(s.conv_root), t.c, t.h, t.w, plan.split, st, cpad, plan.c8_s2d[s.conv_root]);
        else launch_c8_to_planar(c8ptr(s.conv_root), t.p, t.c, t.h, t.w, plan.split, st, cpad, plan.c8_s2d[s.conv_root]);
        return 0;
    }
    if (s.kind == 1) {
        const Layer& L = net.layers[s.layer];
        const DeviceWeights& W = (*dwp_)[s.layer];
        const Tensor& x = plan.blobs[L.bottoms[0]];
        const Tensor& o = plan.blobs[s.out_blob];
        TcConvArgs a;
        memset(&a, 0, sizeof a);
        a.wpk = (const __half*)W.wpk;
        a.bias = W.biasN;
        a.H = o.h; a.W = o.w; a.Cin = W.cinp; a.Cout = L.geti(0, 0); a.N = W.tcN;
        a.s2 = W.tc_s2;
        a.k5 = W.tc_k5;
        if (L.type == "Deconvolution") { a.H = x.h; a.W = x.w; }
        a.split_in = plan.split;
        a.num_sms = num_sms;
        const bool isconv = L.type == "Convolution";
        if (isconv) {
            a.epi = TC_EPI_C8;
            a.out = c8ptr(s.out_blob);
            a.out_plane = (size_t)((o.c + 15) / 16 * 16) * o.h * o.w;  // C8 storage is allocated and converted in multiples of 16 channels
            a.out_s2d = plan.c8_s2d[plan.root[s.out_blob]];
            a.split_out = plan.split;
            if (s.fused_add_blob >= 0) {
                a.res = c8ptr(s.fused_add_blob);
                a.res_plane = (size_t)((plan.blobs[s.fused_add_blob].c + 15) / 16 * 16) * o.h * o.w;
                a.res_split = plan.split;
                a.res_mode = 1;
            }
            if (L.geti(9, 0) == 2) {
                const ParamVal* ap = L.get(10);
                a.act_mode = 1;
                a.slope = ap && !ap->af.empty() ? ap->af[0] : 0.f;
            }
        } else {
            a.epi = TC_EPI_DECONV;
            a.out_f32 = o.p;
            a.ocs = W.ocs;
            a.ps = s.fused_ps_layer >= 0 ? 2 : 1;
            a.act_mode = L.geti(9, 0) == 4 ? 3 : 0;
        }
        if (s.fused_act_layer >= 0) {
            const Layer& A = net.layers[s.fused_act_layer];
            if (A.type == "ReLU") { a.act_mode = 1; a.slope = A.getf(0, 0.f); }
            else if (A.slope.size() == 1) { a.act_mode = 1; a.slope = A.slope[0]; }
            else { a.act_mode = 2; a.prelu = (*dwp_)[s.fused_act_layer].slope; }
        }
        // output-channel slices (W.nchunks > 1): each launch writes W.tcN GEMM columns' worth of channels of the same tensor
        const int cchunk = isconv ? W.tcN : W.ocs;  // channels per slice
        for (int ch = 0; ch < W.nchunks; ch++) {
            TcConvArgs b = a;
            if (W.nchunks > 1) {
                b.wpk = a.wpk + (size_t)ch * W.chunk_elems;
                b.bias = a.bias + (size_t)ch * W.tcN;
                b.Cout = std::min(cchunk, L.geti(0, 0) - ch * cchunk);  // the last slice may be narrower
                if (a.prelu) b.prelu = a.prelu + (size_t)ch * cchunk;
                if (isconv) {
                    const size_t hw = (size_t)o.h * o.w, cg_stride = a.out_s2d ? hw / 4 : hw;
                    b.out_cgroups = o.c / 8;
                    b.out = a.out + (size_t)ch * (cchunk / 8) * cg_stride * 8;
                    if (a.res) b.res = a.res + (size_t)ch * (cchunk / 8) * hw * 8;
                } else {
                    b.out_f32 = a.out_f32 + (size_t)ch * cchunk * o.h * o.w;
                }
            }
            int r = launch_tc_conv(b, c8ptr(L.bottoms[0]), st);
            if (r) { err = "launch_tc_conv failed for " + L.name; return -32; }
        }
        return 0;
    }
    const Layer& L = net.layers[s.layer];
    const DeviceWeights& W = (*dwp_)[s.layer];
    const std::string& T = L.type;
    if (T == "Split" || T == "Crop") return 0;
    auto in = [&](int i) -> const Tensor& { return plan.blobs[L.bottoms[i]]; };
    const Tensor& o = plan.blobs[s.out_blob];
    if (T == "Convolution" || T == "Deconvolution") {
        ConvArgs a;
        memset(&a, 0, sizeof a);
        a.in = in(0).p;
        a.wT = W.wT;
        a.bias = W.bias;
        a.out = o.p;
        a.Cin = in(0).c; a.H = in(0).h; a.W = in(0).w;
        a.Cout = o.c; a.OH = o.h; a.OW = o.w;
        a.ocpad = W.ocpad;
        a.act = L.geti(9, 0);
        const ParamVal* ap = L.get(10);
        if (ap && ap->af.size() > 0) a.act_p0 = ap->af[0];
        if (ap && ap->af.size() > 1) a.act_p1 = ap->af[1];
        if (s.fused_add_blob >= 0) a.res = plan.blobs[s.fused_add_blob].p;
        if (s.fused_act_layer >= 0) {
            const Layer& A = net.layers[s.fused_act_layer];
            if (A.type == "ReLU") {
                float slope = A.getf(0, 0.f);
                a.post_act = 2;
                a.post_p0 = slope;
            } else {  // PReLU
                if (A.slope.size() == 1) { a.post_act = 2; a.post_p0 = A.slope[0]; }
                else { a.post_act = 5; a.post_slope = (*dwp_)[s.fused_act_layer].slope; }
            }
        }
        if (T == "Convolution") {
            int k = L.geti(1, 0), sd = L.geti(3, 1), pad = L.geti(4, 0);
            a.DH = o.h; a.DW = o.w;
            a.in_off_y = a.in_off_x = -pad;
            a.out_mul = 1;
            a.nparity = 1;
            launch_conv(a, k, sd, st);
        } else {
            a.DH = in(0).h; a.DW = in(0).w;
            a.in_off_y = a.in_off_x = -1;
            a.out_mul = 2;
            a.nparity = 4;
            launch_conv(a, 2, 1, st);
        }
    } else if (T == "Concat") {
        size_t off = 0;
        for (size_t i = 0; i < L.bottoms.size(); i++) {
            size_t n = in(i).count();
            cudaMemcpyAsync(o.p + off, in(i).p, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
            off += n;
        }
    } else if (T == "Interp") {
        if (o.h == in(0).h && o.w == in(0).w) cudaMemcpyAsync(o.p, in(0).p, o.count() * sizeof(float), cudaMemcpyDeviceToDevice, st);
        else launch_interp_bilinear(in(0).p, in(0).c, in(0).h, in(0).w, o.p, o.h, o.w, st);
    } else if (T == "PixelShuffle") {
        launch_pixelshuffle(in(0).p, in(0).c, in(0).h, in(0).w, o.p, L.geti(0, 1), st);
    } else if (T == "BinaryOp") {
        int op = L.geti(0, 0);
        if (L.bottoms.size() == 1 || L.geti(1, 0)) {
            static const int map[] = {U_ADD_S, U_SUB_S, U_MUL_S, U_DIV_S, -1, -1, -1, U_RSUB_S, U_RDIV_S};
            if (op < 0 || op > 8 || map[op] < 0) { err = "unsupported scalar BinaryOp in " + L.name; return -30; }
            launch_unary(in(0).p, o.p, o.count(), map[op], L.getf(2, 0.f), 0.f, st);
        } else {
            const Tensor& a = in(0);
            const Tensor& b = in(1);
            auto chan = [](const Tensor& t) { return t.dims == 1 ? t.w : t.c; };
            auto hw = [](const Tensor& t) { return t.dims == 1 ? (size_t)1 : (size_t)t.h * t.w; };
            int oc = chan(o);
            size_t ohw = hw(o);
            bool ok = (chan(a) == oc || chan(a) == 1) && (chan(b) == oc || chan(b) == 1) && (hw(a) == ohw || hw(a) == 1) && (hw(b) == ohw || hw(b) == 1);
            static const int map[] = {B_ADD, B_SUB, B_MUL, B_DIV, B_MAX, B_MIN, B_POW, B_RSUB, B_RDIV};
            if (!ok || op < 0 || op > 8) { err = "unsupported BinaryOp broadcast in " + L.name; return -30; }
            launch_binary(a.p, chan(a), hw(a), b.p, chan(b), hw(b), o.p, oc, ohw, map[op], st);
        }
    } else if (T == "Eltwise") {
        const ParamVal* cf = L.get(1);
        float c0 = 1.f, c1 = 1.f;
        if (cf && cf->af.size() == 2) { c0 = cf->af[0]; c1 = cf->af[1]; }
        launch_eltwise_sum2(in(0).p, in(1).p, c0, c1, o.p, o.count(), st);
    } else if (T == "ReLU") {
        float slope = L.getf(0, 0.f);
        launch_unary(in(0).p, o.p, o.count(), slope == 0.f ? U_RELU : U_LEAKY, slope, 0.f, st);
    } else if (T == "PReLU") {
        int c = in(0).dims == 1 ? in(0).w : in(0).c;
        size_t hw = in(0).dims == 1 ? 1 : (size_t)in(0).h * in(0).w;
        launch_prelu(in(0).p, W.slope, (int)L.slope.size(), o.p, c, hw, st);
    } else if (T == "Sigmoid") {
        launch_unary(in(0).p, o.p, o.count(), U_SIGMOID, 0.f, 0.f, st);
    } else if (T == "Clip") {
        launch_unary(in(0).p, o.p, o.count(), U_CLIP, L.getf(0, -3.4e38f), L.getf(1, 3.4e38f), st);
    } else if (T == "UnaryOp") {
        if (L.geti(0, 0) != 1) { err = "unsupported UnaryOp in " + L.name; return -30; }
        launch_unary(in(0).p, o.p, o.count(), U_NEG, 0.f, 0.f, st);
    } else if (T == "rife.Warp") {
        launch_warp(in(0).p, in(1).p, o.p, in(0).c, in(0).h, in(0).w, st);
    } else if (T == "Pooling") {
        if (in(0).c > pool_scratch_c_) {
            cudaFree(pool_scratch_);
            pool_scratch_ = nullptr;
            pool_scratch_c_ = 0;
            if (cudaMalloc(&pool_scratch_, (size_t)global_avgpool_scratch_floats(in(0).c) * sizeof(float)) == cudaSuccess) pool_scratch_c_ = in(0).c;
            else cudaGetLastError();
        }
        launch_global_avgpool(in(0).p, o.p, in(0).c, (size_t)in(0).h * in(0).w, st, pool_scratch_);
    } else if (T == "InnerProduct") {
        const ParamVal* ap = L.get(10);
        launch_innerproduct(in(0).p, W.wT, W.bias, o.p, (int)in(0).count(), o.w, L.geti(9, 0), ap && !ap->af.empty() ? ap->af[0] : 0.f, st);
    } else {
        err = "no kernel for layer type " + T;
        return -31;
    }
    return 0;
}

}  // namespace rife
