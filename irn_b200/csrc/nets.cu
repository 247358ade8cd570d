// C2/C3/I1/I2: ResNet-50 trunk, CAM head and IRNet edge/displacement heads as a native plan.
//
// Reference: net/resnet50.py:17-91 (Bottleneck / ResNet, strides (2,2,2,1)), net/resnet50_cam.py:55-70
// (CAM.forward), net/resnet50_irn.py:23-133,216-234 (heads, MeanShift, EdgeDisplacement.forward).
//
// The plan owns the repacked weights on the device: FixedBatchNorm folded into (weight, bias), weights
// transposed to [kh*kw*Cin][Cout] for the SIMT kernel.  Activations are NHWC fp32 in a caller-provided
// workspace.  The host walks the fixed topology and enqueues kernels on the caller's stream; nothing
// synchronises.
#include <cuda_fp16.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.h"
#include "conv_simt.cuh"
#include "conv_tc.cuh"
#include "conv_f16.cuh"

namespace irn {

struct Conv {
    int cin = 0, cout = 0, k = 1, stride = 1, pad = 0;
    float* wt = nullptr;     // device [k*k*cin][cout]   (SIMT kernel)
    float* bias = nullptr;   // device [cout] or null
    // tensor-core path: weights [cout][k*k*cin] split into tf32 hi / lo parts
    float* w_hi = nullptr;
    float* w_lo = nullptr;
    int bn = 0;              // N tile (64 or 128); 0 = not eligible
    CUtensorMap map_bhi, map_blo;        // box {32, bn}
    CUtensorMap map_bhi64, map_blo64;    // box {32, 64} (short-K configuration)
    // f16x3 path (conv_f16.cuh): weights [cout][k*k*cin], pre-scaled per output channel by a power of two and split into fp16
    // hi / lo parts, boxes {64 k, 64 | 128 rows}; oscale[cout] = the inverse scale applied in the epilogue
    uint16_t* wb_hi = nullptr;
    uint16_t* wb_lo = nullptr;
    float* oscale = nullptr;
    bool bf_ok = false;
    CUtensorMap map_bf_hi[2], map_bf_lo[2];   // N tile 64, 128 (the latter only when cout allows)
    std::vector<float> host_wt, host_bias;   // folded fp32 weights [K][cout] / bias, kept on the host until the plan is complete
    bool stem_tc = false;    // 7x7/s2 stem repacked as 7 k-blocks of (8 taps x 4 channels) over a zero-haloed NHWC4 input
};

struct Head {            // conv1x1 (no bias) -> GroupNorm(groups) -> [upsample] -> ReLU
    Conv conv;
    int groups = 1, up = 1;
    float* gamma = nullptr;
    float* beta = nullptr;
};

struct Block {
    Conv c1, c2, c3, ds;
    bool has_ds = false;
    // f16x3 mode: conv3 and the projection shortcut as ONE K-concatenated 1x1 conv, out = relu([W3 | Wds] . [t2 ; x_strided] + b3 + bds):
    // the shortcut tensor (as wide as the block's output) is neither written nor read back
    Conv c3ds;
    bool has_c3ds = false;
};

}  // namespace irn

struct irn_net {
    int kind = 0;   // 0 = CAM, 1 = IRN (EdgeDisplacement)
    int conv_mode = 2;   // 0 = SIMT exact-fp32 convolutions only, 1 = tcgen05 3xTF32 where eligible, 2 = tcgen05 f16x3 (default; 3xTF32 / SIMT for the layers it cannot take)
    irn::Conv stem;
    irn::Conv stem_f16;            // the stem repacked for the f16x3 kernel (K = 256 over the NHWC4 halo layout)
    std::vector<irn::Block> blocks[4];
    // CAM
    float* classifier = nullptr;   // [20][2048] (SIMT head kernel)
    irn::Conv cls_conv;            // the same weights as a 2048 -> 64 1x1 conv (rows 20..63 zero) for the tensor-core path
    // IRN
    irn::Head edge[5], dp[7];
    float* edge6_w = nullptr;      // [1][160]
    float* edge6_b = nullptr;      // [1]
    float* dp7_w = nullptr;        // [2][256]
    float* mean_shift = nullptr;   // [2]
    std::vector<void*> allocs;
};

namespace irn {

static const int kPlanes[4] = {64, 128, 256, 512};
static const int kBlocks[4] = {3, 4, 6, 3};
static const int kStrides[4] = {1, 2, 2, 1};   // layer1..4 under the reference's strides=(2,2,2,1)

struct Reader {
    const float* p;
    size_t left;
    bool ok = true;
    const float* take(size_t n) {
        if (n > left) {
            ok = false;
            return nullptr;
        }
        const float* r = p;
        p += n;
        left -= n;
        return r;
    }
};

static int upload(irn_net* net, const std::vector<float>& h, float** out) {
    void* d = nullptr;
    IRN_CUDA(cudaMalloc(&d, h.size() * sizeof(float)));
    net->allocs.push_back(d);
    IRN_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    *out = (float*)d;
    return kOk;
}

static int make_bf16_weights(irn_net* net, Conv& c, const std::vector<float>& wt);

// Reads conv weight [cout][cin][k][k] (+ optional BN gamma, beta, mean, var) and uploads the folded,
// transposed tensors.  Fold: y = (conv - mean) / sqrt(var + 1e-5) * gamma + beta  (net/resnet50.py:11-14).
static int read_conv(irn_net* net, Reader& rd, Conv& c, int cin, int cout, int k, int stride, int pad, bool bn) {
    c.cin = cin; c.cout = cout; c.k = k; c.stride = stride; c.pad = pad;
    const size_t nw = (size_t)cout * cin * k * k;
    const float* w = rd.take(nw);
    const float *ga = nullptr, *be = nullptr, *mu = nullptr, *va = nullptr;
    if (bn) {
        ga = rd.take(cout); be = rd.take(cout); mu = rd.take(cout); va = rd.take(cout);
    }
    if (!rd.ok) return fail(kBadArg, "parameter blob too short (conv %dx%d k%d)", cin, cout, k);
    std::vector<float> wt(nw), bias;
    std::vector<double> scale(cout, 1.0);
    if (bn) {
        bias.resize(cout);
        for (int o = 0; o < cout; ++o) {
            scale[o] = (double)ga[o] / std::sqrt((double)va[o] + 1e-5);
            bias[o] = (float)((double)be[o] - (double)mu[o] * scale[o]);
        }
    }
    for (int o = 0; o < cout; ++o)
        for (int ci = 0; ci < cin; ++ci)
            for (int r = 0; r < k; ++r)
                for (int s = 0; s < k; ++s)
                    wt[((size_t)(r * k + s) * cin + ci) * cout + o] = (float)((double)w[(((size_t)o * cin + ci) * k + r) * k + s] * scale[o]);
    int rc = upload(net, wt, &c.wt);
    if (rc) return rc;
    if (bn && (rc = upload(net, bias, &c.bias))) return rc;
    c.host_wt = wt;
    c.host_bias = bias;
    if ((rc = make_bf16_weights(net, c, wt))) return rc;
    // tensor-core eligibility: 32-channel k slices, 64/128-wide N tiles, 1x1 or 3x3, stride 1 or 2
    c.bn = (cout % 128 == 0) ? 128 : (cout % 64 == 0 ? 64 : 0);
    if (cin % kTcBK != 0 || !(k == 1 || k == 3) || !(stride == 1 || stride == 2)) c.bn = 0;
    if (c.bn) {
        const size_t K = (size_t)k * k * cin;
        std::vector<float> hi(nw), lo(nw);
        for (int o = 0; o < cout; ++o)
            for (size_t kk = 0; kk < K; ++kk) {
                const float v = wt[kk * cout + o];
                uint32_t u;
                std::memcpy(&u, &v, 4);
                // round-to-nearest (ties away) to 10 explicit mantissa bits, like cvt.rna.tf32.f32
                uint32_t h = (u + 0x1000u) & 0xFFFFE000u;
                float hf;
                std::memcpy(&hf, &h, 4);
                if (!std::isfinite(hf)) hf = v;
                hi[(size_t)o * K + kk] = hf;
                lo[(size_t)o * K + kk] = v - hf;
            }
        if ((rc = upload(net, hi, &c.w_hi))) return rc;
        if ((rc = upload(net, lo, &c.w_lo))) return rc;
        const uint64_t dims[2] = {(uint64_t)K, (uint64_t)cout};
        const uint64_t strides[1] = {(uint64_t)K * sizeof(float)};
        const uint32_t box[2] = {(uint32_t)kTcBK, (uint32_t)c.bn};
        if ((rc = make_tensor_map(&c.map_bhi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, c.w_hi, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
        if ((rc = make_tensor_map(&c.map_blo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, c.w_lo, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
        const uint32_t box64[2] = {(uint32_t)kTcBK, 64};
        if ((rc = make_tensor_map(&c.map_bhi64, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, c.w_hi, dims, strides, box64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
        if ((rc = make_tensor_map(&c.map_blo64, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, c.w_lo, dims, strides, box64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    }
    return kOk;
}

// fp16 hi / lo planes of the folded weights, [cout][K] K-major, + tensor maps for 64- / 128-row tiles.  Every output
// channel is first scaled by a power of two so that max |w| lies in [1,2): the lo parts stay clear of fp16's subnormal range
// whatever FixedBatchNorm's gamma / sqrt(var) did to the channel, and the epilogue undoes the scale exactly.
static int make_bf16_weights(irn_net* net, Conv& c, const std::vector<float>& wt /* [K][cout] */) {
    const size_t K = (size_t)c.k * c.k * c.cin;
    c.bf_ok = c.cin % kBfBK == 0 && c.cout % 64 == 0 && (c.k == 1 || c.k == 3) && (c.stride == 1 || c.stride == 2);
    if (!c.bf_ok) return kOk;
    std::vector<uint16_t> hi((size_t)c.cout * K), lo((size_t)c.cout * K);
    std::vector<float> inv(c.cout, 1.0f);
    for (int o = 0; o < c.cout; ++o) {
        float mx = 0.f;
        for (size_t kk = 0; kk < K; ++kk) mx = std::max(mx, std::fabs(wt[kk * c.cout + o]));
        int e = 1;
        if (mx > 0.f && std::isfinite(mx)) std::frexp(mx, &e);          // mx = m * 2^e, m in [0.5, 1)
        const int sh = 1 - e;                                           // w * 2^sh has its maximum in [1, 2)
        inv[o] = std::ldexp(1.0f, -sh);
        for (size_t kk = 0; kk < K; ++kk) {
            const float v = std::ldexp(wt[kk * c.cout + o], sh);
            const __half h = __float2half_rn(v);
            const __half l = __float2half_rn(v - __half2float(h));
            hi[(size_t)o * K + kk] = __half_as_ushort(h);
            lo[(size_t)o * K + kk] = __half_as_ushort(l);
        }
    }
    int rc = upload(net, inv, &c.oscale);
    if (rc) return rc;
    for (int part = 0; part < 2; ++part) {
        void* d = nullptr;
        IRN_CUDA(cudaMalloc(&d, hi.size() * sizeof(uint16_t)));
        net->allocs.push_back(d);
        IRN_CUDA(cudaMemcpy(d, part == 0 ? hi.data() : lo.data(), hi.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        (part == 0 ? c.wb_hi : c.wb_lo) = (uint16_t*)d;
    }
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)c.cout};
    const uint64_t strides[1] = {(uint64_t)K * sizeof(uint16_t)};
    for (int i = 0; i < 2; ++i) {
        const uint32_t rows = 64u << i;
        if (c.cout % rows != 0) continue;
        const uint32_t box[2] = {(uint32_t)kBfBK, rows};
        if ((rc = make_tensor_map(&c.map_bf_hi[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, c.wb_hi, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
        if ((rc = make_tensor_map(&c.map_bf_lo[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, c.wb_lo, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    }
    return kOk;
}

static int read_vec(irn_net* net, Reader& rd, size_t n, float** out) {
    const float* p = rd.take(n);
    if (!rd.ok) return fail(kBadArg, "parameter blob too short (vector of %zu)", n);
    return upload(net, std::vector<float>(p, p + n), out);
}

static int read_trunk(irn_net* net, Reader& rd) {
    int rc = read_conv(net, rd, net->stem, 3, 64, 7, 2, 3, true);
    if (rc) return rc;
    {   // tensor-core stem: K = 7 rows x (8 taps x 4 channels) = 224, tap 7 and channel 3 carry zero weights
        Conv& c = net->stem;
        std::vector<float> host((size_t)49 * 3 * 64);
        IRN_CUDA(cudaMemcpy(host.data(), c.wt, host.size() * sizeof(float), cudaMemcpyDeviceToHost));
        const size_t K = 224;
        std::vector<float> hi(64 * K, 0.f), lo(64 * K, 0.f);
        for (int o = 0; o < 64; ++o)
            for (int r = 0; r < 7; ++r)
                for (int t = 0; t < 7; ++t)
                    for (int ci = 0; ci < 3; ++ci) {
                        const float v = host[((size_t)(r * 7 + t) * 3 + ci) * 64 + o];
                        uint32_t u;
                        std::memcpy(&u, &v, 4);
                        uint32_t h = (u + 0x1000u) & 0xFFFFE000u;
                        float hf;
                        std::memcpy(&hf, &h, 4);
                        if (!std::isfinite(hf)) hf = v;
                        hi[(size_t)o * K + r * 32 + t * 4 + ci] = hf;
                        lo[(size_t)o * K + r * 32 + t * 4 + ci] = v - hf;
                    }
        if ((rc = upload(net, hi, &c.w_hi))) return rc;
        if ((rc = upload(net, lo, &c.w_lo))) return rc;
        const uint64_t dims[2] = {K, 64};
        const uint64_t strides[1] = {K * sizeof(float)};
        const uint32_t box[2] = {(uint32_t)kTcBK, 64};
        if ((rc = make_tensor_map(&c.map_bhi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, c.w_hi, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
        if ((rc = make_tensor_map(&c.map_blo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, c.w_lo, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
        c.stem_tc = true;
        // f16x3 stem: the same rows, K = 8 filter rows x 32 (row 7 zero) = 256 = four 64-wide k-blocks
        Conv& f = net->stem_f16;
        f.cin = 256; f.cout = 64; f.k = 1; f.stride = 1; f.pad = 0;
        std::vector<float> wt((size_t)256 * 64, 0.f);
        for (int o = 0; o < 64; ++o)
            for (int r = 0; r < 7; ++r)
                for (int t = 0; t < 7; ++t)
                    for (int ci = 0; ci < 3; ++ci)
                        wt[(size_t)(r * 32 + t * 4 + ci) * 64 + o] = host[((size_t)(r * 7 + t) * 3 + ci) * 64 + o];
        f.bias = c.bias;
        if ((rc = make_bf16_weights(net, f, wt))) return rc;
    }
    int cin = 64;
    for (int l = 0; l < 4; ++l) {
        net->blocks[l].resize(kBlocks[l]);
        for (int b = 0; b < kBlocks[l]; ++b) {
            Block& blk = net->blocks[l][b];
            const int planes = kPlanes[l], stride = b == 0 ? kStrides[l] : 1;
            if ((rc = read_conv(net, rd, blk.c1, cin, planes, 1, 1, 0, true))) return rc;
            if ((rc = read_conv(net, rd, blk.c2, planes, planes, 3, stride, 1, true))) return rc;   // stride on conv2 (net/resnet50.py:24)
            if ((rc = read_conv(net, rd, blk.c3, planes, planes * 4, 1, 1, 0, true))) return rc;
            blk.has_ds = b == 0;
            if (blk.has_ds && (rc = read_conv(net, rd, blk.ds, cin, planes * 4, 1, stride, 0, true))) return rc;
            if (blk.has_ds && planes % kBfBK == 0 && cin % kBfBK == 0) {
                Conv& f = blk.c3ds;
                const int cout = planes * 4, K3 = planes, Kd = cin;
                f.cin = K3 + Kd; f.cout = cout; f.k = 1; f.stride = 1; f.pad = 0;
                std::vector<float> wt((size_t)(K3 + Kd) * cout), bias(cout);
                std::copy(blk.c3.host_wt.begin(), blk.c3.host_wt.end(), wt.begin());
                std::copy(blk.ds.host_wt.begin(), blk.ds.host_wt.end(), wt.begin() + (size_t)K3 * cout);
                for (int o = 0; o < cout; ++o) bias[o] = blk.c3.host_bias[o] + blk.ds.host_bias[o];
                if ((rc = upload(net, bias, &f.bias))) return rc;
                if ((rc = make_bf16_weights(net, f, wt))) return rc;
                blk.has_c3ds = f.bf_ok;
            }
            cin = planes * 4;
        }
    }
    // the host copies were only needed to build the fused convs
    net->stem.host_wt.clear(); net->stem.host_wt.shrink_to_fit();
    for (int l = 0; l < 4; ++l)
        for (Block& blk : net->blocks[l])
            for (Conv* c : {&blk.c1, &blk.c2, &blk.c3, &blk.ds}) {
                std::vector<float>().swap(c->host_wt);
                std::vector<float>().swap(c->host_bias);
            }
    return kOk;
}

static int read_head(irn_net* net, Reader& rd, Head& h, int cin, int cout, int groups, int up) {
    int rc = read_conv(net, rd, h.conv, cin, cout, 1, 1, 0, false);
    if (rc) return rc;
    h.groups = groups;
    h.up = up;
    if ((rc = read_vec(net, rd, cout, &h.gamma))) return rc;
    return read_vec(net, rd, cout, &h.beta);
}

// ------------------------------------------------------------------ launch helpers
static inline int conv_out(int n, int k, int s, int p) { return (n + 2 * p - k) / s + 1; }

template <int BN, int STAGES, int NACC>
static int launch_tc(const Conv& c, const float* in, int B, int H, int W, int Ho, int Wo, const float* residual, float* out, bool relu,
                     cudaStream_t st) {
    static DeviceOnce once;
    const int ds = once.slot();
    if (once.need(ds)) {
        IRN_CUDA(cudaFuncSetAttribute((conv_tc_kernel<BN, STAGES, NACC>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem_bytes(BN, STAGES)));
        once.done[ds] = true;
    }
    TcMaps maps;
    maps.b_hi = BN == 64 && c.bn == 128 ? c.map_bhi64 : c.map_bhi;
    maps.b_lo = BN == 64 && c.bn == 128 ? c.map_blo64 : c.map_blo;
    const uint64_t dims[4] = {(uint64_t)c.cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)c.cin * 4, (uint64_t)W * c.cin * 4, (uint64_t)H * W * c.cin * 4};
    const uint32_t box[4] = {(uint32_t)kTcBK, (uint32_t)(kTcTW * c.stride), (uint32_t)(kTcTH * c.stride), 1};
    const uint32_t estr[4] = {1, (uint32_t)c.stride, (uint32_t)c.stride, 1};
    int rc = make_tensor_map(&maps.a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, in, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, estr);
    if (rc) return rc;
    TcArgs a;
    a.bias = c.bias; a.residual = residual; a.out = out;
    a.B = B; a.Ho = Ho; a.Wo = Wo; a.Cout = c.cout; a.Cin = c.cin; a.ksize = c.k; a.stride = c.stride; a.pad = c.pad;
    a.relu = relu ? 1 : 0;
    a.tiles_x = (Wo + kTcTW - 1) / kTcTW;
    a.tiles_y = (Ho + kTcTH - 1) / kTcTH;
    a.mode = 0;
    dim3 grid((unsigned)(a.tiles_x * a.tiles_y * B * (c.cout / BN)));
    conv_tc_kernel<BN, STAGES, NACC><<<grid, kTcThreads, tc_smem_bytes(BN, STAGES), st>>>(maps, a);
    IRN_LAUNCH_CHECK("conv_tc_kernel");
    return kOk;
}

// Tensor-core stem: x4 = zero-haloed NHWC4 input [B, Hin+6, Win+8, 4]; out NHWC [B,Ho,Wo,64] with bias + ReLU.
// A-from-TMEM persistent kernel for the 64-channel layers (IRN_TC_PERSIST_TS=0 selects the shared-memory-operand one for A/B runs)
static bool persist_ts_enabled() {
    static const int v = getenv("IRN_TC_PERSIST_TS") ? atoi(getenv("IRN_TC_PERSIST_TS")) : 1;
    return v != 0;
}
static int persist_ts_attr() {
    static DeviceOnce once;
    const int ds = once.slot();
    if (once.need(ds)) {
        IRN_CUDA(cudaFuncSetAttribute(conv_tc_persist_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPtsSmem));
        once.done[ds] = true;
    }
    return kOk;
}

static int launch_tc_stem(const Conv& c, const float* x4, int B, int Hin, int Win, float* out, cudaStream_t st) {
    static DeviceOnce once;
    const int ds = once.slot();
    if (once.need(ds)) {
        IRN_CUDA(cudaFuncSetAttribute(conv_tc_persist_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcPersistCfg<64>::kSmem));
        int dev = 0;
        IRN_CUDA(cudaGetDevice(&dev));
        IRN_CUDA(cudaDeviceGetAttribute(&once.n_sm[ds], cudaDevAttrMultiProcessorCount, dev));
        once.done[ds] = true;
    }
    const int n_sm = once.n_sm[ds];
    const int Hp = Hin + 6, Wp = Win + 8;
    const int Ho = conv_out(Hin, 7, 2, 3), Wo = conv_out(Win, 7, 2, 3);
    TcMaps maps;
    maps.b_hi = c.map_bhi;
    maps.b_lo = c.map_blo;
    // dim0: the 32 contiguous floats (8 px x 4 ch) of one filter-row window; dim1: output column (windows overlap: stride 2 px = 32 B);
    // dim2: padded input row; dim3: image
    const uint64_t dims[4] = {32, (uint64_t)Wo, (uint64_t)Hp, (uint64_t)B};
    const uint64_t strides[3] = {32, (uint64_t)Wp * 16, (uint64_t)Hp * Wp * 16};
    const uint32_t box[4] = {32, (uint32_t)kTcTW, (uint32_t)(kTcTH * 2), 1};
    const uint32_t estr[4] = {1, 1, 2, 1};
    int rc = make_tensor_map(&maps.a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, estr);
    if (rc) return rc;
    TcArgs a;
    a.bias = c.bias; a.residual = nullptr; a.out = out;
    a.B = B; a.Ho = Ho; a.Wo = Wo; a.Cout = 64; a.Cin = 32; a.ksize = 7; a.stride = 2; a.pad = 3; a.relu = 1;
    a.tiles_x = (Wo + kTcTW - 1) / kTcTW;
    a.tiles_y = (Ho + kTcTH - 1) / kTcTH;
    a.mode = 1;
    const long long total = (long long)a.tiles_x * a.tiles_y * B;
    if (persist_ts_enabled()) {
        int rc2 = persist_ts_attr();
        if (rc2) return rc2;
        conv_tc_persist_ts_kernel<<<(unsigned)(total < n_sm ? total : n_sm), kTcPersistThreads, kPtsSmem, st>>>(maps, a);
        IRN_LAUNCH_CHECK("conv_tc_persist_ts_kernel(stem)");
        return kOk;
    }
    conv_tc_persist_kernel<64><<<(unsigned)(total < n_sm ? total : n_sm), kTcPersistThreads, TcPersistCfg<64>::kSmem, st>>>(maps, a);
    IRN_LAUNCH_CHECK("conv_tc_persist_kernel(stem)");
    return kOk;
}

template <int BN>
static int launch_tc_persist(const Conv& c, const float* in, int B, int H, int W, int Ho, int Wo, const float* residual, float* out,
                             bool relu, cudaStream_t st) {
    static DeviceOnce once;
    const int ds = once.slot();
    if (once.need(ds)) {
        IRN_CUDA(cudaFuncSetAttribute(conv_tc_persist_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TcPersistCfg<BN>::kSmem));
        int dev = 0;
        IRN_CUDA(cudaGetDevice(&dev));
        IRN_CUDA(cudaDeviceGetAttribute(&once.n_sm[ds], cudaDevAttrMultiProcessorCount, dev));
        once.done[ds] = true;
    }
    const int n_sm = once.n_sm[ds];
    TcMaps maps;
    maps.b_hi = BN == 64 && c.bn == 128 ? c.map_bhi64 : c.map_bhi;
    maps.b_lo = BN == 64 && c.bn == 128 ? c.map_blo64 : c.map_blo;
    const uint64_t dims[4] = {(uint64_t)c.cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)c.cin * 4, (uint64_t)W * c.cin * 4, (uint64_t)H * W * c.cin * 4};
    const uint32_t box[4] = {(uint32_t)kTcBK, (uint32_t)(kTcTW * c.stride), (uint32_t)(kTcTH * c.stride), 1};
    const uint32_t estr[4] = {1, (uint32_t)c.stride, (uint32_t)c.stride, 1};
    int rc = make_tensor_map(&maps.a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, in, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, estr);
    if (rc) return rc;
    TcArgs a;
    a.bias = c.bias; a.residual = residual; a.out = out;
    a.B = B; a.Ho = Ho; a.Wo = Wo; a.Cout = c.cout; a.Cin = c.cin; a.ksize = c.k; a.stride = c.stride; a.pad = c.pad;
    a.relu = relu ? 1 : 0;
    a.tiles_x = (Wo + kTcTW - 1) / kTcTW;
    a.tiles_y = (Ho + kTcTH - 1) / kTcTH;
    a.mode = 0;
    const long long total = (long long)a.tiles_x * a.tiles_y * B * (c.cout / BN);
    const unsigned grid = (unsigned)(total < n_sm ? total : n_sm);
    if (BN == 64 && persist_ts_enabled()) {
        if ((rc = persist_ts_attr())) return rc;
        conv_tc_persist_ts_kernel<<<grid, kTcPersistThreads, kPtsSmem, st>>>(maps, a);
        IRN_LAUNCH_CHECK("conv_tc_persist_ts_kernel");
        return kOk;
    }
    conv_tc_persist_kernel<BN><<<grid, kTcPersistThreads, TcPersistCfg<BN>::kSmem, st>>>(maps, a);
    IRN_LAUNCH_CHECK("conv_tc_persist_kernel");
    return kOk;
}

static int launch_tc_ts(const Conv& c, const float* in, int B, int H, int W, int Ho, int Wo, const float* residual, float* out, bool relu,
                        cudaStream_t st) {
    static DeviceOnce once;
    const int ds = once.slot();
    if (once.need(ds)) {
        IRN_CUDA(cudaFuncSetAttribute(conv_tc_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTsSmem));
        once.done[ds] = true;
    }
    TcMaps maps;
    maps.b_hi = c.map_bhi;
    maps.b_lo = c.map_blo;
    const uint64_t dims[4] = {(uint64_t)c.cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)c.cin * 4, (uint64_t)W * c.cin * 4, (uint64_t)H * W * c.cin * 4};
    const uint32_t box[4] = {(uint32_t)kTcBK, (uint32_t)(kTcTW * c.stride), (uint32_t)(kTcTH * c.stride), 1};
    const uint32_t estr[4] = {1, (uint32_t)c.stride, (uint32_t)c.stride, 1};
    int rc = make_tensor_map(&maps.a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, in, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, estr);
    if (rc) return rc;
    TcArgs a;
    a.bias = c.bias; a.residual = residual; a.out = out;
    a.B = B; a.Ho = Ho; a.Wo = Wo; a.Cout = c.cout; a.Cin = c.cin; a.ksize = c.k; a.stride = c.stride; a.pad = c.pad;
    a.relu = relu ? 1 : 0;
    a.tiles_x = (Wo + kTcTW - 1) / kTcTW;
    a.tiles_y = (Ho + kTcTH - 1) / kTcTH;
    a.mode = 0;
#ifdef IRN_EXPERIMENTAL
    static const int use_pair = getenv("IRN_TC_PAIR") ? atoi(getenv("IRN_TC_PAIR")) : 0;   // measured: no gain (1322 vs 1323 us on the 3x3x512 layer), kept for A/B
    if (use_pair) {
        static DeviceOnce once2;
        const int ds2 = once2.slot();
        if (once2.need(ds2)) {
            IRN_CUDA(cudaFuncSetAttribute(conv_tc_ts2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTsSmem));
            once2.done[ds2] = true;
        }
        maps.b_hi = c.map_bhi64;    // each CTA of a pair loads 64 of the 128 weight rows and multicasts them
        maps.b_lo = c.map_blo64;
        const long long m_tiles = (long long)a.tiles_x * a.tiles_y * B;
        dim3 grid2((unsigned)(((m_tiles + 1) / 2) * (c.cout / 128) * 2));
        conv_tc_ts2_kernel<<<grid2, kTsThreads, kTsSmem, st>>>(maps, a);
        IRN_LAUNCH_CHECK("conv_tc_ts2_kernel");
        return kOk;
    }
#endif
    dim3 grid((unsigned)(a.tiles_x * a.tiles_y * B * (c.cout / 128)));
    conv_tc_ts_kernel<<<grid, kTsThreads, kTsSmem, st>>>(maps, a);
    IRN_LAUNCH_CHECK("conv_tc_ts_kernel");
    return kOk;
}

// ---- f16x3 kernels (conv_f16.cuh)
static int f16_mode_flags() {
    static const int spin = getenv("IRN_F16_SPIN") ? atoi(getenv("IRN_F16_SPIN")) : 1;     // 0: suspending try_wait on the critical path (A/B runs)
    static const int pf = getenv("IRN_F16_RES_PREFETCH") ? atoi(getenv("IRN_F16_RES_PREFETCH")) : 1;   // 0: no L2 prefetch of the next tile's residual (A/B runs)
    return (spin ? 0 : 4) | (pf ? 0 : 8);
}

// Second input of a K-concatenated 1x1 conv (Block::c3ds): NHWC [B, H2, W2, cin2] sampled with pixel stride `stride`
struct F16Second {
    const float* in = nullptr;
    int H = 0, W = 0, cin = 0, stride = 1;
};

template <int BN, int NACC, int NSLOT, bool HALO>
static int launch_f16(const Conv& c, const float* in, int B, int H, int W, int Ho, int Wo, const float* residual, float* out, bool relu,
                      cudaStream_t st, const F16Second* second = nullptr) {
    using Cfg = F16Cfg<BN, NACC, NSLOT>;
    constexpr size_t smem = HALO ? Cfg::kSmemHalo : Cfg::kSmem;
    static DeviceOnce once;
    const int ds = once.slot();
    if (once.need(ds)) {
        if (HALO)
            IRN_CUDA(cudaFuncSetAttribute((conv_f16_halo_kernel<BN, NACC, NSLOT>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else
            IRN_CUDA(cudaFuncSetAttribute((conv_f16_kernel<BN, NACC, NSLOT>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int dev = 0;
        IRN_CUDA(cudaGetDevice(&dev));
        IRN_CUDA(cudaDeviceGetAttribute(&once.n_sm[ds], cudaDevAttrMultiProcessorCount, dev));
        once.done[ds] = true;
    }
    const int mi = BN == 64 ? 0 : 1;
    TcMaps maps;
    maps.b_hi = c.map_bf_hi[mi];
    maps.b_lo = c.map_bf_lo[mi];
    const int cin1 = second ? c.cin - second->cin : c.cin;        // channels of the first input
    const uint64_t dims[4] = {(uint64_t)cin1, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t strides[3] = {(uint64_t)cin1 * 4, (uint64_t)W * cin1 * 4, (uint64_t)H * W * cin1 * 4};
    const uint32_t box[4] = {32u, (uint32_t)(HALO ? kHaloW : kTcTW * c.stride), (uint32_t)(HALO ? kHaloH : kTcTH * c.stride), 1};
    const uint32_t estr[4] = {1, (uint32_t)(HALO ? 1 : c.stride), (uint32_t)(HALO ? 1 : c.stride), 1};
    int rc = make_tensor_map(&maps.a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, in, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, estr);
    if (rc) return rc;
    maps.a2 = maps.a;
    if (second) {
        if (HALO || c.k != 1) return fail(kUnsupported, "launch_f16: a second input needs the plain 1x1 kernel");
        const uint64_t d2[4] = {(uint64_t)second->cin, (uint64_t)second->W, (uint64_t)second->H, (uint64_t)B};
        const uint64_t s2[3] = {(uint64_t)second->cin * 4, (uint64_t)second->W * second->cin * 4, (uint64_t)second->H * second->W * second->cin * 4};
        const uint32_t b2[4] = {32u, (uint32_t)(kTcTW * second->stride), (uint32_t)(kTcTH * second->stride), 1};
        const uint32_t e2[4] = {1, (uint32_t)second->stride, (uint32_t)second->stride, 1};
        if ((rc = make_tensor_map(&maps.a2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, second->in, d2, s2, b2, CU_TENSOR_MAP_SWIZZLE_128B, e2))) return rc;
    }
    TcArgs a;
    a.bias = c.bias; a.residual = residual; a.out = out;
    a.oscale = c.oscale;
    a.B = B; a.Ho = Ho; a.Wo = Wo; a.Cout = c.cout; a.Cin = c.cin; a.ksize = c.k; a.stride = c.stride; a.pad = c.pad;
    a.relu = relu ? 1 : 0;
    a.tiles_x = (Wo + kTcTW - 1) / kTcTW;
    a.tiles_y = (Ho + kTcTH - 1) / kTcTH;
    a.mode = f16_mode_flags();
    if (second) {
        a.kb_split = cin1 / kBfBK;
        a.stride2 = second->stride;
    }
    const long long total = (long long)a.tiles_x * a.tiles_y * B * (c.cout / BN);
    const int n_sm = once.n_sm[ds];
    const unsigned grid = (unsigned)(total < n_sm ? total : n_sm);
    if (HALO) {
        conv_f16_halo_kernel<BN, NACC, NSLOT><<<grid, kBfThreads, smem, st>>>(maps, a);
        IRN_LAUNCH_CHECK("conv_f16_halo_kernel");
    } else {
        conv_f16_kernel<BN, NACC, NSLOT><<<grid, kBfThreads, smem, st>>>(maps, a);
        IRN_LAUNCH_CHECK("conv_f16_kernel");
    }
    return kOk;
}

// 7x7 / stride-2 stem on the f16x3 kernel: x4 = zero-haloed NHWC4 input [B, Hin+6, Win+8, 4] (as for launch_tc_stem)
static int launch_f16_stem(const Conv& c, const float* x4, int B, int Hin, int Win, float* out, cudaStream_t st) {
    using Cfg = F16Cfg<64, 1, 4>;
    static DeviceOnce once;
    const int ds = once.slot();
    if (once.need(ds)) {
        IRN_CUDA(cudaFuncSetAttribute((conv_f16_kernel<64, 1, 4>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem));
        int dev = 0;
        IRN_CUDA(cudaGetDevice(&dev));
        IRN_CUDA(cudaDeviceGetAttribute(&once.n_sm[ds], cudaDevAttrMultiProcessorCount, dev));
        once.done[ds] = true;
    }
    const int Hp = Hin + 6, Wp = Win + 8;
    const int Ho = conv_out(Hin, 7, 2, 3), Wo = conv_out(Win, 7, 2, 3);
    TcMaps maps;
    maps.b_hi = c.map_bf_hi[0];
    maps.b_lo = c.map_bf_lo[0];
    // dim0: the 32 contiguous floats (8 px x 4 ch) of one filter-row window; dim1: output column (windows overlap: stride 2 px = 32 B);
    // dim2: padded input row; dim3: image
    const uint64_t dims[4] = {32, (uint64_t)Wo, (uint64_t)Hp, (uint64_t)B};
    const uint64_t strides[3] = {32, (uint64_t)Wp * 16, (uint64_t)Hp * Wp * 16};
    const uint32_t box[4] = {32, (uint32_t)kTcTW, (uint32_t)(kTcTH * 2), 1};
    const uint32_t estr[4] = {1, 1, 2, 1};
    int rc = make_tensor_map(&maps.a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, estr);
    if (rc) return rc;
    maps.a2 = maps.a;
    TcArgs a;
    a.bias = c.bias; a.residual = nullptr; a.out = out;
    a.oscale = c.oscale;
    a.B = B; a.Ho = Ho; a.Wo = Wo; a.Cout = 64; a.Cin = 256; a.ksize = 1; a.stride = 2; a.pad = 0; a.relu = 1;
    a.tiles_x = (Wo + kTcTW - 1) / kTcTW;
    a.tiles_y = (Ho + kTcTH - 1) / kTcTH;
    a.mode = f16_mode_flags() | 1;
    const long long total = (long long)a.tiles_x * a.tiles_y * B;
    const int n_sm = once.n_sm[ds];
    conv_f16_kernel<64, 1, 4><<<(unsigned)(total < n_sm ? total : n_sm), kBfThreads, Cfg::kSmem, st>>>(maps, a);
    IRN_LAUNCH_CHECK("conv_f16_kernel(stem)");
    return kOk;
}

// f16x3 dispatch.  Reductions with K >= 512 keep the hi*hi and the cross terms in separate TMEM accumulators (the tensor core's
// accumulate truncates: conv_f16.cuh, f16_issue3), shorter ones use one accumulator per tile and two accumulator sets so that the
// epilogue of a tile overlaps the next tile's mainloop (they are memory-bound).  3x3 / stride 1 convs take the halo-tile kernel.
template <bool HALO>
static int dispatch_f16(const Conv& c, const float* in, int B, int H, int W, int Ho, int Wo, const float* residual, float* out, bool relu,
                        cudaStream_t st, const F16Second* second = nullptr) {
    const int K = c.k * c.k * c.cin;
    static const int acc_min_k = getenv("IRN_F16_ACC_MINK") ? atoi(getenv("IRN_F16_ACC_MINK")) : 512;
    const bool sep = K >= acc_min_k;
    if (c.cout % 128 == 0) {
        if (sep) return launch_f16<128, 2, 4, HALO>(c, in, B, H, W, Ho, Wo, residual, out, relu, st, second);
        return launch_f16<128, 1, 4, HALO>(c, in, B, H, W, Ho, Wo, residual, out, relu, st, second);
    }
    if (sep) return launch_f16<64, 2, 4, HALO>(c, in, B, H, W, Ho, Wo, residual, out, relu, st, second);
    return launch_f16<64, 1, 4, HALO>(c, in, B, H, W, Ho, Wo, residual, out, relu, st, second);
}

static int run_conv_bf16(const Conv& c, const float* in, int B, int H, int W, int Ho, int Wo, const float* residual, float* out, bool relu,
                         cudaStream_t st) {
    static const int use_halo = getenv("IRN_F16_HALO") ? atoi(getenv("IRN_F16_HALO")) : 1;
    if (use_halo && c.k == 3 && c.stride == 1 && c.pad == 1) return dispatch_f16<true>(c, in, B, H, W, Ho, Wo, residual, out, relu, st);
    return dispatch_f16<false>(c, in, B, H, W, Ho, Wo, residual, out, relu, st);
}

static int run_conv(const irn_net* net, const Conv& c, const float* in, int B, int H, int W, const float* residual, float* out, bool relu,
                    cudaStream_t st, int* Ho_, int* Wo_) {
    ConvGeom g;
    g.B = B; g.H = H; g.W = W; g.Cin = c.cin;
    g.Ho = conv_out(H, c.k, c.stride, c.pad);
    g.Wo = conv_out(W, c.k, c.stride, c.pad);
    g.Cout = c.cout; g.k = c.k; g.stride = c.stride; g.pad = c.pad;
    if (Ho_) *Ho_ = g.Ho;
    if (Wo_) *Wo_ = g.Wo;
    if (net->conv_mode == 2 && c.bf_ok) return run_conv_bf16(c, in, B, H, W, g.Ho, g.Wo, residual, out, relu, st);
    if (net->conv_mode >= 1 && c.bn) {
        // short reductions (K <= 256) are latency-bound per tile: 64-wide tiles with a 2-stage pipeline and two TMEM
        // accumulators let two CTAs share an SM; long reductions use the 3-stage, 3-accumulator configuration
        const int K = c.k * c.k * c.cin;
        static const int persist_max_k = getenv("IRN_TC_PERSIST_MAXK") ? atoi(getenv("IRN_TC_PERSIST_MAXK")) : 640;
        if (K <= persist_max_k) {   // persistent, epilogue-overlapped kernel for the short reductions
            if (c.bn == 128) return launch_tc_persist<128>(c, in, B, H, W, g.Ho, g.Wo, residual, out, relu, st);
            return launch_tc_persist<64>(c, in, B, H, W, g.Ho, g.Wo, residual, out, relu, st);
        }
        if (K <= 128) return launch_tc<64, 2, 2>(c, in, B, H, W, g.Ho, g.Wo, residual, out, relu, st);
        if (K <= 256) return launch_tc<64, 2, 3>(c, in, B, H, W, g.Ho, g.Wo, residual, out, relu, st);
        static const int use_ts = getenv("IRN_TC_TS") ? atoi(getenv("IRN_TC_TS")) : 1;
        if (c.bn == 128 && use_ts) return launch_tc_ts(c, in, B, H, W, g.Ho, g.Wo, residual, out, relu, st);   // A operand from TMEM
        if (c.bn == 128) return launch_tc<128, 3, 3>(c, in, B, H, W, g.Ho, g.Wo, residual, out, relu, st);
        return launch_tc<64, 3, 3>(c, in, B, H, W, g.Ho, g.Wo, residual, out, relu, st);
    }
    const int M = B * g.Ho * g.Wo;
    dim3 grid((M + kBM - 1) / kBM, (c.cout + kBN - 1) / kBN);
    if (c.cin % 16 == 0)
        conv_simt_kernel<true><<<grid, 256, 0, st>>>(in, c.wt, c.bias, residual, out, g, relu ? 1 : 0);
    else
        conv_simt_kernel<false><<<grid, 256, 0, st>>>(in, c.wt, c.bias, residual, out, g, relu ? 1 : 0);
    IRN_LAUNCH_CHECK("conv_simt_kernel");
    return kOk;
}

struct Arena {
    char* base;
    size_t size, off = 0;
    bool ok = true;
    float* take(size_t n_floats) {
        const size_t bytes = align_up(n_floats * sizeof(float), 256);
        if (off + bytes > size) {
            ok = false;
            return nullptr;
        }
        float* p = (float*)(base + off);
        off += bytes;
        return p;
    }
};

struct TrunkShapes {
    int H1, W1, H2, W2;          // after stem conv, after maxpool (= layer1 grid)
    int Hl[4], Wl[4];            // output grid of layer1..4
    size_t max_act;              // largest activation (floats) among stem out / block tensors
};

static TrunkShapes trunk_shapes(int B, int H, int W) {
    TrunkShapes s;
    s.H1 = conv_out(H, 7, 2, 3); s.W1 = conv_out(W, 7, 2, 3);
    s.H2 = conv_out(s.H1, 3, 2, 1); s.W2 = conv_out(s.W1, 3, 2, 1);
    int h = s.H2, w = s.W2;
    s.max_act = (size_t)B * s.H1 * s.W1 * 64;
    for (int l = 0; l < 4; ++l) {
        // conv1 of the first block still runs on the incoming grid with `planes` channels
        s.max_act = std::max(s.max_act, (size_t)B * h * w * kPlanes[l]);
        h = conv_out(h, 3, kStrides[l], 1);
        w = conv_out(w, 3, kStrides[l], 1);
        s.Hl[l] = h; s.Wl[l] = w;
        s.max_act = std::max(s.max_act, (size_t)B * h * w * kPlanes[l] * 4);
    }
    return s;
}

// Runs input layout transform + stem + maxpool + layer1..4 on x_nchw [B,3,H,W] zero-padded (logically) to Hin x Win.
// feats[0] = after maxpool (x1 of IRNet), feats[1..4] = layer outputs.  When `keep` is set every feats[i] lives in
// its own arena buffer (IRNet taps them); otherwise buffers rotate.
static int run_trunk(const irn_net* net, const float* x_nchw, int B, int H, int W, int Hin, int Win, Arena& ar, bool keep,
                     const float* feats[5], TrunkShapes& sh, cudaStream_t st) {
    sh = trunk_shapes(B, Hin, Win);
    const bool stem_tc = net->conv_mode >= 1 && net->stem.stem_tc;
    float* x_in = stem_tc ? ar.take((size_t)B * (Hin + 6) * (Win + 8) * 4) : ar.take((size_t)B * Hin * Win * 3);
    float* stem_out = ar.take((size_t)B * sh.H1 * sh.W1 * 64);
    float* pool_out = ar.take((size_t)B * sh.H2 * sh.W2 * 64);
    float* t1 = ar.take(sh.max_act);
    float* t2 = ar.take(sh.max_act);
    float* dsb = ar.take(sh.max_act);
    float* ping[2] = {ar.take(sh.max_act), ar.take(sh.max_act)};
    if (!ar.ok) return fail(kWorkspace, "network workspace too small");
    int rc;
    if (stem_tc) {
        const size_t total = (size_t)B * (Hin + 6) * (Win + 8);
        nchw_to_nhwc4_halo_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x_nchw, (float4*)x_in, B, H, W, Hin + 6, Win + 8);
        IRN_LAUNCH_CHECK("nchw_to_nhwc4_halo_kernel");
        static const int stem_f16 = getenv("IRN_F16_STEM") ? atoi(getenv("IRN_F16_STEM")) : 1;
        if (net->conv_mode == 2 && net->stem_f16.bf_ok && stem_f16) {
            if ((rc = launch_f16_stem(net->stem_f16, x_in, B, Hin, Win, stem_out, st))) return rc;
        } else if ((rc = launch_tc_stem(net->stem, x_in, B, Hin, Win, stem_out, st))) return rc;
    } else {
        const size_t total = (size_t)B * Hin * Win * 3;
        nchw_to_nhwc_pad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x_nchw, x_in, B, 3, H, W, Hin, Win);
        IRN_LAUNCH_CHECK("nchw_to_nhwc_pad_kernel");
        if ((rc = run_conv(net, net->stem, x_in, B, Hin, Win, nullptr, stem_out, true, st, nullptr, nullptr))) return rc;
    }
    {
        const size_t total = (size_t)B * sh.H2 * sh.W2 * 16;
        maxpool3s2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(stem_out, pool_out, B, sh.H1, sh.W1, 64, sh.H2, sh.W2);
        IRN_LAUNCH_CHECK("maxpool3s2_kernel");
    }
    feats[0] = pool_out;
    const float* x = pool_out;
    int h = sh.H2, w = sh.W2, flip = 0;
    for (int l = 0; l < 4; ++l) {
        const int nb = (int)net->blocks[l].size();
        for (int b = 0; b < nb; ++b) {
            const Block& blk = net->blocks[l][b];
            int ho, wo;
            if ((rc = run_conv(net, blk.c1, x, B, h, w, nullptr, t1, true, st, nullptr, nullptr))) return rc;
            if ((rc = run_conv(net, blk.c2, t1, B, h, w, nullptr, t2, true, st, &ho, &wo))) return rc;
            static const int fuse_ds = getenv("IRN_F16_FUSE_DS") ? atoi(getenv("IRN_F16_FUSE_DS")) : 1;
            const bool fused = net->conv_mode == 2 && blk.has_c3ds && fuse_ds;
            const float* res = x;
            if (blk.has_ds && !fused) {
                if ((rc = run_conv(net, blk.ds, x, B, h, w, nullptr, dsb, false, st, nullptr, nullptr))) return rc;
                res = dsb;
            }
            float* out;
            if (keep && b == nb - 1) {   // IRNet taps every stage output: give it a buffer of its own
                out = ar.take((size_t)B * ho * wo * blk.c3.cout);
                if (!ar.ok) return fail(kWorkspace, "network workspace too small");
            } else {
                out = ping[flip];
                flip ^= 1;
            }
            if (fused) {   // conv3 + projection shortcut in one reduction: [t2 ; x sampled at the block's stride]
                F16Second sec;
                sec.in = x; sec.H = h; sec.W = w; sec.cin = blk.ds.cin; sec.stride = blk.ds.stride;
                if ((rc = dispatch_f16<false>(blk.c3ds, t2, B, ho, wo, ho, wo, nullptr, out, true, st, &sec))) return rc;
            } else if ((rc = run_conv(net, blk.c3, t2, B, ho, wo, res, out, true, st, nullptr, nullptr))) return rc;   // out += residual; relu (net/resnet50.py:51-52)
            x = out;
            h = ho;
            w = wo;
        }
        feats[l + 1] = x;
    }
    return kOk;
}

static size_t trunk_workspace_floats(int B, int H, int W, bool keep) {
    TrunkShapes s = trunk_shapes(B, H, W);
    size_t n = (size_t)B * (H + 6) * (W + 8) * 4 + (size_t)B * s.H1 * s.W1 * 64 + (size_t)B * s.H2 * s.W2 * 64 + 5 * s.max_act;
    if (keep)
        for (int l = 0; l < 4; ++l) n += (size_t)B * s.Hl[l] * s.Wl[l] * kPlanes[l] * 4;
    return n + 64 * 32;   // alignment slack (256 B per buffer)
}

}  // namespace irn

using namespace irn;

// ---- single convolution as a plan of its own (unit tests / integration of other networks)
struct irn_conv {
    irn_net holder;   // owns the device allocations
    irn::Conv conv;
};

extern "C" int irn_conv_create(const float* weight_oihw, const float* bn4 /* gamma,beta,mean,var or NULL */, int cin, int cout, int k,
                               int stride, int pad, irn_conv** out) {
    if (!weight_oihw || !out || cin <= 0 || cout <= 0 || k <= 0 || stride <= 0 || pad < 0) return fail(kBadArg, "irn_conv_create: bad argument");
    const size_t nw = (size_t)cout * cin * k * k;
    std::vector<float> blob(weight_oihw, weight_oihw + nw);
    if (bn4) blob.insert(blob.end(), bn4, bn4 + 4 * (size_t)cout);
    irn_conv* c = new irn_conv();
    Reader rd{blob.data(), blob.size()};
    int rc = read_conv(&c->holder, rd, c->conv, cin, cout, k, stride, pad, bn4 != nullptr);
    if (rc) {
        for (void* p : c->holder.allocs) cudaFree(p);
        delete c;
        return rc;
    }
    *out = c;
    return kOk;
}

extern "C" void irn_conv_destroy(irn_conv* c) {
    if (!c) return;
    for (void* p : c->holder.allocs) cudaFree(p);
    delete c;
}

// in NHWC fp32 [B,H,W,cin] -> out NHWC [B,Ho,Wo,cout]; residual NHWC like out or NULL; mode as irn_net_set_conv_mode
extern "C" int irn_conv_forward(irn_conv* c, const float* in, int B, int H, int W, const float* residual, float* out, int relu, int mode,
                                irn_stream_t stream) {
    launch_counter() = 0;
    if (!c || !in || !out || B <= 0 || H <= 0 || W <= 0) return fail(kBadArg, "irn_conv_forward: bad argument");
    if (mode < 0 || mode > 2) return fail(kBadArg, "irn_conv_forward: mode must be 0, 1 or 2");
    if (mode == 1 && c->conv.bn == 0) return fail(kUnsupported, "irn_conv_forward: this convolution is not eligible for the tensor-core kernel (Cin %% 32, Cout %% 64, k in {1,3}, stride in {1,2})");
    if (mode == 2 && !c->conv.bf_ok) return fail(kUnsupported, "irn_conv_forward: this convolution is not eligible for the f16x3 kernel (Cin %% 64, Cout %% 64, k in {1,3}, stride in {1,2})");
    c->holder.conv_mode = mode;
    return run_conv(&c->holder, c->conv, in, B, H, W, residual, out, relu != 0, (cudaStream_t)stream, nullptr, nullptr);
}

extern "C" int irn_net_set_conv_mode(irn_net* net, int mode) {
    if (!net || mode < 0 || mode > 2) return fail(kBadArg, "irn_net_set_conv_mode: mode must be 0 (SIMT fp32), 1 (tcgen05 3xTF32) or 2 (tcgen05 f16x3)");
    net->conv_mode = mode;
    return kOk;
}

extern "C" int irn_net_get_conv_mode(const irn_net* net) { return net ? net->conv_mode : -1; }

extern "C" void irn_net_destroy(irn_net* net) {
    if (!net) return;
    for (void* p : net->allocs) cudaFree(p);
    delete net;
}

extern "C" int irn_cam_net_create(const float* params, size_t n_floats, irn_net** out) {
    if (!params || !out) return fail(kBadArg, "irn_cam_net_create: null pointer");
    irn_net* net = new irn_net();
    net->kind = 0;
    Reader rd{params, n_floats};
    int rc = read_trunk(net, rd);
    if (!rc) {
        const float* cw = rd.p;
        rc = read_vec(net, rd, (size_t)20 * 2048, &net->classifier);
        if (!rc) {
            std::vector<float> padded((size_t)64 * 2048, 0.f);
            std::memcpy(padded.data(), cw, (size_t)20 * 2048 * sizeof(float));
            Reader r2{padded.data(), padded.size()};
            rc = read_conv(net, r2, net->cls_conv, 2048, 64, 1, 1, 0, false);
        }
    }
    if (!rc && rd.left != 0) rc = fail(kBadArg, "irn_cam_net_create: %zu unread floats in the parameter blob", rd.left);
    if (rc) {
        irn_net_destroy(net);
        return rc;
    }
    *out = net;
    return kOk;
}

extern "C" int irn_irn_net_create(const float* params, size_t n_floats, irn_net** out) {
    if (!params || !out) return fail(kBadArg, "irn_irn_net_create: null pointer");
    irn_net* net = new irn_net();
    net->kind = 1;
    Reader rd{params, n_floats};
    int rc = read_trunk(net, rd);
    // heads in the order of net/resnet50_irn.py:23-93
    static const int e_cin[5] = {64, 256, 512, 1024, 2048}, e_up[5] = {1, 1, 2, 4, 4};
    for (int i = 0; i < 5 && !rc; ++i) rc = read_head(net, rd, net->edge[i], e_cin[i], 32, 4, e_up[i]);
    if (!rc) rc = read_vec(net, rd, 160, &net->edge6_w);
    if (!rc) rc = read_vec(net, rd, 1, &net->edge6_b);
    static const int d_cin[7] = {64, 256, 512, 1024, 2048, 768, 448}, d_cout[7] = {64, 128, 256, 256, 256, 256, 256},
                     d_g[7] = {8, 16, 16, 16, 16, 16, 16}, d_up[7] = {1, 1, 1, 2, 2, 2, 1};
    for (int i = 0; i < 7 && !rc; ++i) rc = read_head(net, rd, net->dp[i], d_cin[i], d_cout[i], d_g[i], d_up[i]);
    if (!rc) rc = read_vec(net, rd, 2 * 256, &net->dp7_w);
    if (!rc) rc = read_vec(net, rd, 2, &net->mean_shift);
    if (!rc && rd.left != 0) rc = fail(kBadArg, "irn_irn_net_create: %zu unread floats in the parameter blob", rd.left);
    if (rc) {
        irn_net_destroy(net);
        return rc;
    }
    *out = net;
    return kOk;
}

extern "C" size_t irn_cam_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || (B & 1) || H <= 0 || W <= 0) return 0;
    const TrunkShapes sh = trunk_shapes(B, H, W);
    return (trunk_workspace_floats(B, H, W, false) + (size_t)B * sh.Hl[3] * sh.Wl[3] * 64 + 128) * sizeof(float);
}

// CAM.forward for P = B/2 (image, flipped image) pairs: x NCHW fp32 [B,3,H,W] -> cam [P,20,ceil(H/16),ceil(W/16)]
extern "C" int irn_cam_forward(const irn_net* net, const float* x_nchw, int B, int H, int W, float* cam_out, void* workspace,
                               size_t workspace_bytes, irn_stream_t stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    launch_counter() = 0;
    if (!net || net->kind != 0 || !x_nchw || !cam_out || !workspace) return fail(kBadArg, "irn_cam_forward: bad argument");
    if (B <= 0 || (B & 1) || H <= 0 || W <= 0) return fail(kBadArg, "irn_cam_forward: B must be a positive even number (image + flipped image), got B=%d H=%d W=%d", B, H, W);
    if (((uintptr_t)workspace & 255) != 0) return fail(kBadArg, "irn_cam_forward: workspace must be 256-byte aligned");
    Arena ar{(char*)workspace, workspace_bytes};
    const float* feats[5];
    TrunkShapes sh;
    int rc = run_trunk(net, x_nchw, B, H, W, H, W, ar, false, feats, sh, st);
    if (rc) return rc;
    const int P = B / 2, h = sh.Hl[3], w = sh.Wl[3];
    if (net->conv_mode >= 1 && net->cls_conv.bn) {
        // classifier as a 2048 -> 64 tensor-core conv with fused ReLU (the one-warp-per-pixel head re-reads the 160 KB weight
        // matrix per pixel), then flip-add + NHWC -> NCHW on the 20 real channels
        float* tmp = ar.take((size_t)B * h * w * 64);
        if (!ar.ok) return fail(kWorkspace, "irn_cam_forward: workspace too small");
        if ((rc = run_conv(net, net->cls_conv, feats[4], B, h, w, nullptr, tmp, true, st, nullptr, nullptr))) return rc;
        const size_t total = (size_t)P * 20 * h * w;
        cam_flip_add_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(tmp, cam_out, P, h, w, 64);
        IRN_LAUNCH_CHECK("cam_flip_add_kernel");
        return kOk;
    }
    const size_t warps = (size_t)P * h * w;
    cam_head_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(feats[4], net->classifier, cam_out, P, h, w, 2048);
    IRN_LAUNCH_CHECK("cam_head_kernel");
    return kOk;
}

static size_t irn_head_floats(int B, const TrunkShapes& s) {
    const size_t g2 = (size_t)B * s.Hl[0] * s.Wl[0];       // stride-4 grid (x1, x2)
    const size_t g3 = (size_t)B * s.Hl[1] * s.Wl[1];       // stride-8 grid (x3)
    size_t n = 0;
    n += g2 * 256;                 // raw conv output scratch (largest head conv: 256 ch on the stride-4 grid)
    n += g2 * 160;                 // edge concat
    n += g3 * 768;                 // dp3|dp4|dp5 concat
    n += g2 * 448;                 // dp1|dp2|dp_up3 concat
    n += g2 * 256;                 // dp7 activations
    n += g2 * 3;                   // edge logits + dp
    n += (size_t)B * 16 * 4 + 64;  // GN sums (fp64): B x (<= 16 groups, enforced in run_head) x (sum, sum of squares)
    return n + 64 * 16;
}

extern "C" size_t irn_edge_displacement_workspace_bytes(int P, int H, int W, int crop_size) {
    if (P <= 0 || H <= 0 || W <= 0 || H > crop_size || W > crop_size) return 0;
    const int B = 2 * P;
    TrunkShapes s = trunk_shapes(B, crop_size, crop_size);
    return (trunk_workspace_floats(B, crop_size, crop_size, true) + irn_head_floats(B, s) + 64) * sizeof(float);
}

static int run_head(const irn_net* net, const Head& hd, const float* x, int B, int H, int W, float* raw, float* stats, float* dst, int Hd, int Wd, int Cd,
                    int coff, cudaStream_t st) {
    int rc = run_conv(net, hd.conv, x, B, H, W, nullptr, raw, false, st, nullptr, nullptr);
    if (rc) return rc;
    if (hd.conv.cout > 256 || 256 % hd.conv.cout != 0 || hd.groups > 16) return fail(kUnsupported, "GroupNorm head with %d channels / %d groups", hd.conv.cout, hd.groups);
    IRN_CUDA(cudaMemsetAsync(stats, 0, (size_t)B * hd.groups * 2 * sizeof(double), st));
    {
        const int slices = std::max(1, std::min(256, (H * W) / 256));
        gn_partial_kernel<<<dim3(slices, B), 256, 0, st>>>(raw, (double*)stats, H * W, hd.conv.cout, hd.groups);
        IRN_LAUNCH_CHECK("gn_partial_kernel");
    }
    const size_t total = (size_t)B * Hd * Wd * (hd.conv.cout / 4);
    gn_up_relu_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(raw, (const double*)stats, hd.gamma, hd.beta, dst, B, H, W, hd.conv.cout, hd.groups,
                                                                      hd.up, Hd, Wd, Cd, coff);
    IRN_LAUNCH_CHECK("gn_up_relu_kernel");
    return kOk;
}

// EdgeDisplacement.forward: x NCHW fp32 [2,3,H,W] (image, flipped image) -> edge [1,fh,fw], dp [2,fh,fw],
// fh = ceil(H/4), fw = ceil(W/4).  The input is zero-padded to crop_size x crop_size (net/resnet50_irn.py:226).
extern "C" int irn_edge_displacement_forward(const irn_net* net, const float* x_nchw, int P, int H, int W, int crop_size, float* edge_out,
                                             float* dp_out, void* workspace, size_t workspace_bytes, irn_stream_t stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    launch_counter() = 0;
    if (!net || net->kind != 1 || !x_nchw || !edge_out || !dp_out || !workspace) return fail(kBadArg, "irn_edge_displacement_forward: bad argument");
    if (P <= 0 || H <= 0 || W <= 0 || H > crop_size || W > crop_size || (crop_size % 16) != 0)
        return fail(kBadArg, "irn_edge_displacement_forward: need P > 0 and 0 < H,W <= crop_size (multiple of 16); got P=%d H=%d W=%d crop=%d", P, H, W, crop_size);
    if (((uintptr_t)workspace & 255) != 0) return fail(kBadArg, "irn_edge_displacement_forward: workspace must be 256-byte aligned");
    const int B = 2 * P, S = crop_size;
    Arena ar{(char*)workspace, workspace_bytes};
    const float* f[5];
    TrunkShapes sh;
    int rc = run_trunk(net, x_nchw, B, H, W, S, S, ar, true, f, sh, st);
    if (rc) return rc;
    const int h2 = sh.Hl[0], w2 = sh.Wl[0];   // stride 4 (x1, x2)
    const int h3 = sh.Hl[1], w3 = sh.Wl[1];   // stride 8 (x3)
    const int h4 = sh.Hl[2], w4 = sh.Wl[2];   // stride 16 (x4, x5)
    const size_t g2 = (size_t)B * h2 * w2, g3 = (size_t)B * h3 * w3;
    float* raw = ar.take(g2 * 256);
    float* ecat = ar.take(g2 * 160);
    float* dcat3 = ar.take(g3 * 768);
    float* dcat2 = ar.take(g2 * 448);
    float* dp7a = ar.take(g2 * 256);
    float* elog = ar.take(g2);
    float* dlog = ar.take(g2 * 2);
    float* stats = ar.take((size_t)B * 16 * 4 + 64);
    if (!ar.ok) return fail(kWorkspace, "irn_edge_displacement_forward: workspace too small");

    // edge branch (net/resnet50_irn.py:117-122): every map lands on the stride-4 grid, cropped to edge2's size
    const int eh[5] = {h2, h2, h3, h4, h4}, ew[5] = {w2, w2, w3, w4, w4};
    for (int i = 0; i < 5; ++i)
        if ((rc = run_head(net, net->edge[i], f[i], B, eh[i], ew[i], raw, stats, ecat, h2, w2, 160, 32 * i, st))) return rc;
    conv1x1_smalln_kernel<1><<<(unsigned)((g2 * 32 + 255) / 256), 256, 0, st>>>(ecat, net->edge6_w, net->edge6_b, nullptr, elog, g2, 160);
    IRN_LAUNCH_CHECK("conv1x1_smalln_kernel<1>");

    // displacement branch (net/resnet50_irn.py:124-131)
    if ((rc = run_head(net, net->dp[0], f[0], B, h2, w2, raw, stats, dcat2, h2, w2, 448, 0, st))) return rc;      // dp1 64
    if ((rc = run_head(net, net->dp[1], f[1], B, h2, w2, raw, stats, dcat2, h2, w2, 448, 64, st))) return rc;     // dp2 128
    if ((rc = run_head(net, net->dp[2], f[2], B, h3, w3, raw, stats, dcat3, h3, w3, 768, 0, st))) return rc;      // dp3
    if ((rc = run_head(net, net->dp[3], f[3], B, h4, w4, raw, stats, dcat3, h3, w3, 768, 256, st))) return rc;    // dp4 up x2, crop to dp3
    if ((rc = run_head(net, net->dp[4], f[4], B, h4, w4, raw, stats, dcat3, h3, w3, 768, 512, st))) return rc;    // dp5 up x2
    if ((rc = run_head(net, net->dp[5], dcat3, B, h3, w3, raw, stats, dcat2, h2, w2, 448, 192, st))) return rc;   // dp6 up x2, crop to dp2
    if ((rc = run_head(net, net->dp[6], dcat2, B, h2, w2, raw, stats, dp7a, h2, w2, 256, 0, st))) return rc;      // dp7 conv/GN/ReLU
    conv1x1_smalln_kernel<2><<<(unsigned)((g2 * 32 + 255) / 256), 256, 0, st>>>(dp7a, net->dp7_w, nullptr, net->mean_shift, dlog, g2, 256);
    IRN_LAUNCH_CHECK("conv1x1_smalln_kernel<2>");

    const int fh = (H - 1) / 4 + 1, fw = (W - 1) / 4 + 1;
    edge_dp_tail_kernel<<<dim3((fh * fw + 255) / 256, P), 256, 0, st>>>(elog, dlog, edge_out, dp_out, h2, w2, fh, fw);
    IRN_LAUNCH_CHECK("edge_dp_tail_kernel");
    return kOk;
}
