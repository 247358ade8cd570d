// tcgen05 implicit-GEMM convolution, "f16x3" split arithmetic: NHWC fp32 activations in and out, every product evaluated as
//      a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo        a_hi = fp16(a), a_lo = fp16(a - a_hi)  (likewise w, split on the host)
// with kind::f16 MMAs (fp16 operands, fp32 accumulation in TMEM).  fp16 carries the same 11 significant bits as tf32, so the
// split keeps ~22 bits per operand -- the accuracy class of the 3xTF32 kernels in conv_tc.cuh -- at HALF their tensor time
// (K = 16 per MMA instead of 8) and two thirds of their operand bytes.  (A bf16 split, tried first, keeps only 16 bits and
// buys nothing: profiles/r02_conv_f16x3_accuracy.md.)
// Range: fp16 spans 6e-8 .. 65504.  Weights are pre-scaled per output channel by a power of two (max |w| in [1,2), undone
// exactly in the epilogue); activations beyond +-65504 saturate (ResNet-50 activations stay orders of magnitude below) and
// lose relative precision below 6e-5, where their contribution to a sum is below fp32 resolution anyway.
//
// Two persistent kernels share every piece (1x1 / 3x3, stride 1 / 2, Cin % 64 == 0, Cout % 64 == 0):
//   conv_f16_kernel       one activation tile per (tap, 64-channel slice)
//   conv_f16_halo_kernel  3x3 / stride 1: ONE 10 x 18 pixel halo tile per 64-channel slice, the nine taps read shifted windows
//   GEMM tile   M = 8x16 output pixels (128 TMEM lanes), N = BN in {64, 128} channels, K walked in 64-channel k-blocks per tap
//   warp 0      TMA producer: fp32 activation tiles as 4-D boxes {32 ch, px, rows, 1 image} (zero fill = padding, element strides =
//               stride-2 convs) + the fp16 weight tiles w_hi / w_lo as 2-D boxes {64 k, BN} (SWIZZLE_128B)
//   warps 4-11  split: thread = tile row = TMEM lane, two warps per lane quarter (one per 32-channel half of the k-block); reads its
//               128-byte row slice once, writes packed fp16 A_hi / A_lo (2 k-elements per 32-bit column) straight into one of
//               NSLOT TMEM A slots with tcgen05.st.  (With the MMAs issued back to back -- elect_one -- a lone split warp per
//               scheduler, ~1000 cycles of LDS -> convert -> tcgen05.st -> barrier per k-block, was the next pacer.)
//   warp 1      MMA issuer: per k-block 4 k-steps (K = 16) x 2-3 MMAs (f16_issue_kblock), A from TMEM, B from shared memory
//   warps 12-19 epilogue: TMEM -> swizzled shared staging -> coalesced (x 2^-s, +bias, +residual, ReLU) fp32 stores, 32 channels
//               (128-byte row segments) at a time; with two accumulator sets it overlaps the next tile's mainloop
//   TMEM        columns [0, 512 - 64 NSLOT): SETS x NACC accumulators of BN columns; the top 64 NSLOT columns: A slots (hi 32 | lo 32)
// Accumulators (NACC): the tensor core's fp32 accumulate TRUNCATES; see f16_issue_kblock.
// A slots (NSLOT = 4): the ring  split -> a_ready -> MMA -> commit -> a_free -> split  is covered four k-blocks deep.
// Issue discipline: see elect_one -- the MMA / TMA issue loops run on a converged warp.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "conv_tc.cuh"

namespace irn {

constexpr int kBfBK = 64;            // channels per k-block
constexpr int kBfThreads = 640;      // 20 warps = 5 warpgroups: [TMA, MMA, 2 idle] | 8 split | 8 epilogue (register budgets per group: f16_regs)
constexpr int kHaloW = kTcTW + 2, kHaloH = kTcTH + 2;                 // 18 x 10 pixels
constexpr int kHaloHalfBytes = 23 * 1024;                            // 180 rows x 128 B = 23,040 B, padded to a 1 KB multiple (swizzle atom)
constexpr int kHaloBytes = 2 * kHaloHalfBytes;

// Register budgets per warpgroup (setmaxnreg): the kernel is compiled for 640 threads x 96 registers; the issue warpgroup and the
// split warpgroups hand registers back, the epilogue warpgroups -- which keep a 32-row x 64-channel residual slice in flight per
// thread plus 32-64 accumulator values -- take them: 4 x 32 x 40 + 8 x 32 x 72 + 8 x 32 x 152 = 62,464 <= 640 x 96 = 61,440 + ... see static_assert.
constexpr int kRegsIssue = 40, kRegsSplit = 72, kRegsEpilogue = 144;
static_assert(4 * 32 * kRegsIssue + 8 * 32 * kRegsSplit + 8 * 32 * kRegsEpilogue <= 640 * 96, "register pool");
template <int N, bool INC>
__device__ __forceinline__ void f16_regs() {
    if (INC)
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
    else
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

template <int BN, int NACC, int NSLOT>
struct F16Cfg {
    static constexpr int kAccCols = 512 - 64 * NSLOT;
    static constexpr int kSets = (2 * NACC * BN <= kAccCols) ? 2 : 1;
    static constexpr uint32_t kAcol = (uint32_t)kAccCols;          // first TMEM column of the A slots
    static constexpr int kBBytes = BN * 128;                       // one weight plane: BN rows x 64 fp16
    static constexpr int kStagingBytes = 8 * 32 * 32 * 4;          // 8 epilogue warps x 32 rows x 32 floats
    // plain kernel: stage = A (2 x 32 fp32 channels) | B_hi | B_lo
    static constexpr int kStages = BN == 128 ? 3 : 4;
    static constexpr int kStageBytes = 2 * 16384 + 2 * kBBytes;
    static constexpr size_t kSmem = 1024 + (size_t)kStages * kStageBytes + kStagingBytes + 256;
    // halo kernel: two halo buffers | weight stages (B_hi | B_lo of one tap of one slice)
    static constexpr int kStagesB = BN == 128 ? 3 : 4;
    static constexpr int kBStageBytes = 2 * kBBytes;
    static constexpr size_t kSmemHalo = 1024 + 2 * (size_t)kHaloBytes + (size_t)kStagesB * kBStageBytes + kStagingBytes + 256;
    static_assert(NACC * BN * kSets <= kAccCols, "accumulators overlap the TMEM A slots");
    static_assert(BN == 64 || BN == 128, "N tile 64 or 128");
    static_assert(NACC == 1 || NACC == 2, "one accumulator, or main | cross");
};

// kind::f16 instruction descriptor: D fp32 (1 @4), A and B fp16 (0 @7, 0 @10), both K-major, N>>3 @17, M>>4 @24
__device__ __forceinline__ constexpr uint32_t f16_idesc(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void f16_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// (a, b) -> packed fp16 pair {low half = hi(a), high half = hi(b)} and the packed pair of the residuals.  hi(x) = x TRUNCATED to 11
// significant bits: the round-toward-zero conversion does it while packing (fp16 and tf32 share the mantissa width), one LOP3 per
// element rebuilds the same value in fp32 for the residual x - hi(x), which is exact in fp32 (13 significant bits) and rounded once,
// to 11 bits, by its own conversion: |error| <= 2^-21 |x|.  (Exact packing needs 2^-14 <= |x| <= 65504; beyond that the conversion
// saturates, below it the two parts disagree by < 6e-8 absolute.)  Three ALU ops per element; the first version (round-to-nearest hi
// with two integer ops) had four.
__device__ __forceinline__ void f16_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rz.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));      // upper half <- first source operand
    const float ha = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u), hb = __uint_as_float(__float_as_uint(b) & 0xFFFFE000u);
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - hb), "f"(a - ha));
}

// One 128-byte row slice (32 fp32 channels, SWIZZLE_128B: 16-byte chunk c of row r lives at chunk c ^ (r % 8)) -> 16 packed
// fp16 hi pairs + 16 packed lo pairs in k order.
__device__ __forceinline__ void f16_split_row(const float4* arow, int key, uint32_t* hi, uint32_t* lo) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 v = arow[c ^ key];
        f16_split2(v.x, v.y, hi[c * 2], lo[c * 2]);
        f16_split2(v.z, v.w, hi[c * 2 + 1], lo[c * 2 + 1]);
    }
}

// waits on the per-k-block critical path: spinning poll unless args.mode bit 2 asks for the suspending try_wait (A/B runs)
__device__ __forceinline__ void f16_wait(uint64_t* bar, uint32_t parity, bool spin) {
    if (spin)
        mbar_wait_spin(bar, parity);
    else
        mbar_wait(bar, parity);
}

// One lane of a CONVERGED warp.  tcgen05.mma / tcgen05.commit / cp.async.bulk.tensor are warp-uniform instructions: issued under
// `if (lane == 0)` the compiler wraps every one of them in its own ELECT + BRA.U.ANY loop (~100 cycles of single-thread latency
// per MMA: 12 MMAs = the 1200-1500 cycles per k-block that paced every layer, whatever its N, in round 1 and in this round's
// first f16 kernels); under elect.sync in a converged warp they issue back to back.
__device__ __forceinline__ bool elect_one() { return tc_elect_one(); }

// The MMAs of one k-block (4 k-steps of 16).  The tensor core's fp32 accumulate TRUNCATES (measured on B200: every MMA into an
// accumulator shrinks it by ~1.5e-8 of its value: -1.3e-5 after the 864 MMAs of a K = 4608 reduction,
// profiles/r02_conv_f16x3_accuracy.md), so for long reductions the small cross terms get their own accumulator:
//   NACC = 1   hi*hi, lo*hi, hi*lo into one accumulator (three N = BN MMAs per k-step)
//   NACC = 2   accumulators [main | cross] side by side in TMEM and the weight planes [w_hi | w_lo] side by side in shared memory:
//              ONE N = 2 BN MMA computes A_hi x [w_hi | w_lo] = [hi*hi | hi*lo] into [main | cross], a second N = BN MMA adds
//              A_lo x w_hi to `cross`: same tensor time, two instructions instead of three, and only the hi*hi MMAs truncate the
//              large sum.  The epilogue adds main + cross in IEEE fp32.
template <int BN, int NACC>
__device__ __forceinline__ void f16_issue_kblock(uint32_t acc, uint32_t a_t, uint32_t b_hi, int kb) {
    const uint64_t d_hi = tc_smem_desc(b_hi), d_lo = tc_smem_desc(b_hi + (uint32_t)(BN * 128));
#pragma unroll
    for (int j = 0; j < 4; ++j) {          // UMMA_K = 16 fp16 = 32 bytes inside the 128-byte swizzle atom (+2 in the descriptor) = 8 TMEM columns of A
        const uint32_t ta_hi = a_t + j * 8, ta_lo = a_t + 32 + j * 8;
        if (NACC == 1) {
            constexpr uint32_t idesc = f16_idesc(128, BN);
            f16_mma_ts(acc, ta_hi, d_hi + 2 * j, idesc, (kb | j) != 0);
            f16_mma_ts(acc, ta_lo, d_hi + 2 * j, idesc, 1);
            f16_mma_ts(acc, ta_hi, d_lo + 2 * j, idesc, 1);
        } else {
            f16_mma_ts(acc, ta_hi, d_hi + 2 * j, f16_idesc(128, 2 * BN), (kb | j) != 0);
            f16_mma_ts(acc + BN, ta_lo, d_hi + 2 * j, f16_idesc(128, BN), 1);
        }
    }
}

__device__ __forceinline__ void f16_tile_coords(const TcArgs& args, int n_tiles, int BN, int id, int& b, int& oy0, int& ox0, int& n0) {
    const int m = id / n_tiles;
    n0 = (id % n_tiles) * BN;
    ox0 = (m % args.tiles_x) * kTcTW;
    oy0 = ((m / args.tiles_x) % args.tiles_y) * kTcTH;
    b = m / (args.tiles_x * args.tiles_y);
}

// Epilogue shared by the f16x3 kernels (warps 12..19 of a CTA): TMEM lane quarter = warp % 4, channel half = (warp - 12) / 4; 32
// channels per pass through a 4 KB swizzled staging tile (thread = row on the way in, 8 lanes x float4 per row on the way out).
template <int BN, int NACC, int SETS>
__device__ __forceinline__ void f16_epilogue(const TcArgs& args, float* staging, uint64_t* tmem_full, uint64_t* tmem_empty, uint32_t tmem_base,
                                             int warp, int lane, int total, int n_tiles) {
    const int q = warp & 3;
    const int hf = (warp - 12) >> 2;
    constexpr int kCols = BN / 2;
    constexpr int kChunks = kCols / 32;
    float* stg = staging + (size_t)(warp - 12) * 32 * 32;
    const int sub = lane >> 3, c8 = lane & 7;
    uint32_t ti = 0;
    for (int id = blockIdx.x; id < total; id += gridDim.x, ++ti) {
        int b, oy0, ox0, n0;
        f16_tile_coords(args, n_tiles, BN, id, b, oy0, ox0, n0);
        const uint32_t set = SETS == 2 ? (ti & 1) : 0, use = SETS == 2 ? (ti >> 1) : ti;
        const size_t tile_base = (((size_t)b * args.Ho + oy0) * args.Wo + ox0) * args.Cout + n0 + hf * kCols + c8 * 4;
        // pass i of a 32-row quarter touches tile row q*2 + i/4, columns (i%4)*4 + sub: two strides instead of eight offsets
        const uint32_t row_stride = (uint32_t)args.Wo * (uint32_t)args.Cout, col_stride = 4u * (uint32_t)args.Cout;
        const uint32_t off0 = (uint32_t)(q * 2) * row_stride + (uint32_t)sub * (uint32_t)args.Cout;
        auto off = [&](int i) -> uint32_t { return off0 + (uint32_t)(i >> 2) * row_stride + (uint32_t)(i & 3) * col_stride; };
        uint32_t okmask = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            okmask |= ((oy0 + q * 2 + (i >> 2) < args.Ho && ox0 + (i & 3) * 4 + sub < args.Wo) ? 1u : 0u) << i;
        const float* res_base = args.residual ? args.residual + tile_base : nullptr;
        float* out_base = args.out + tile_base;
        float4 res[kChunks * 8];          // the warp's whole residual slice is in flight before the accumulator is ready
#pragma unroll
        for (int cc = 0; cc < kChunks; ++cc)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                res[cc * 8 + i] = (res_base && ((okmask >> i) & 1u)) ? __ldg(reinterpret_cast<const float4*>(res_base + off(i) + cc * 32))
                                                                    : make_float4(0.f, 0.f, 0.f, 0.f);
        // the NEXT tile's residual slice is asked into L2 now (no registers held): its loads, issued only after this tile's stores,
        // then pay an L2 hit instead of a DRAM round trip
        if (res_base && id + (int)gridDim.x < total && !(args.mode & 8)) {
            int nb, noy0, nox0, nn0;
            f16_tile_coords(args, n_tiles, BN, id + (int)gridDim.x, nb, noy0, nox0, nn0);
            const float* nres = args.residual + (((size_t)nb * args.Ho + noy0) * args.Wo + nox0) * args.Cout + nn0 + hf * kCols + c8 * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (noy0 + q * 2 + (i >> 2) < args.Ho && nox0 + (i & 3) * 4 + sub < args.Wo) {
#pragma unroll
                    for (int cc = 0; cc < kChunks; ++cc)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(nres + off(i) + cc * 32));
                }
        }
        mbar_wait(&tmem_full[set], use & 1);
        tc_fence_after();
#pragma unroll
        for (int cc = 0; cc < kChunks; ++cc) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + set * (uint32_t)(NACC * BN) + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * kCols + cc * 32);
            tc_ld32(taddr, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (NACC >= 2) {
                uint32_t u[32];
                tc_ld32(taddr + (uint32_t)((NACC - 1) * BN), u);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
            }
            if (cc == kChunks - 1) {          // last TMEM read of this tile: hand the accumulator set back
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[set]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)       // row `lane`, float4 slot j, XOR-swizzled: conflict-free both ways
                *reinterpret_cast<uint4*>(stg + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            __syncwarp();
            const int nch = n0 + hf * kCols + cc * 32 + c8 * 4;
            float4 bi = make_float4(0.f, 0.f, 0.f, 0.f), sc = make_float4(1.f, 1.f, 1.f, 1.f);
            if (args.bias) bi = __ldg(reinterpret_cast<const float4*>(args.bias + nch));
            if (args.oscale) sc = __ldg(reinterpret_cast<const float4*>(args.oscale + nch));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = i * 4 + sub;
                float4 o = *reinterpret_cast<const float4*>(stg + r * 32 + ((c8 ^ (r & 7)) << 2));
                if ((okmask >> i) & 1u) {
                    // the scale is a power of two: the product is exact, so the (possibly contracted) multiply-add rounds once like the add alone
                    o.x = o.x * sc.x + bi.x; o.y = o.y * sc.y + bi.y; o.z = o.z * sc.z + bi.z; o.w = o.w * sc.w + bi.w;
                    const float4 rv = res[cc * 8 + i];
                    o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
                    if (args.relu) {
                        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                    }
                    *reinterpret_cast<float4*>(out_base + off(i) + cc * 32) = o;
                }
            }
            __syncwarp();   // the staging tile is rewritten by the next pass
        }
    }
}

// split warp: hi / lo registers of its 32-channel half -> TMEM A slot (hi columns [0,32), lo columns [32,64)), then signal
__device__ __forceinline__ void f16_store_slot(uint32_t taddr, const uint32_t (&hi)[16], const uint32_t (&lo)[16], uint64_t* ready, int lane) {
    tc_st16(taddr, hi);
    tc_st16(taddr + 32, lo);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(ready);
}

template <int BN, int NACC, int NSLOT>
__global__ void __launch_bounds__(kBfThreads, 1)
conv_f16_kernel(const __grid_constant__ TcMaps maps, const TcArgs args) {
    using Cfg = F16Cfg<BN, NACC, NSLOT>;
    constexpr int S = Cfg::kStages;
    constexpr int SETS = Cfg::kSets;
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    unsigned char* smem = tc_smem_raw;
    float* staging = reinterpret_cast<float*>(smem + S * Cfg::kStageBytes);
    uint64_t* bars = (uint64_t*)(smem + S * Cfg::kStageBytes + Cfg::kStagingBytes);
    uint64_t* full = bars;                     // [S] TMA landed
    uint64_t* empty = bars + S;                // [S] MMAs finished reading the stage (they start after the split warps read it)
    uint64_t* a_ready = bars + 2 * S;          // [NSLOT] A_hi / A_lo of a k-block are in the TMEM slot
    uint64_t* a_free = a_ready + NSLOT;        // [NSLOT] MMAs finished reading the TMEM slot
    uint64_t* tmem_full = a_free + NSLOT;      // [2] MMA -> epilogue
    uint64_t* tmem_empty = tmem_full + 2;      // [2] epilogue -> MMA
    uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = args.Cout / BN;
    const int total = args.tiles_x * args.tiles_y * args.B * n_tiles;
    const int cblocks = args.Cin / kBfBK;
    const int KB = args.ksize * args.ksize * cblocks;
    const bool spin = (args.mode & 4) == 0;

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) __trap();
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int i = 0; i < NSLOT; ++i) {
            mbar_init(&a_ready[i], 8);
            mbar_init(&a_free[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 8);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) f16_regs<kRegsIssue, false>();
    if (warp == 0) {
        // ---- TMA producer: the whole warp walks the loop (converged), one elected lane issues
        if (elect_one()) {
            tma_prefetch_desc(&maps.a);
            tma_prefetch_desc(&maps.b_hi);
            tma_prefetch_desc(&maps.b_lo);
            if (args.kb_split < KB) tma_prefetch_desc(&maps.a2);
        }
        uint32_t g = 0;
        for (int id = blockIdx.x; id < total; id += gridDim.x) {
            int b, oy0, ox0, n0;
            f16_tile_coords(args, n_tiles, BN, id, b, oy0, ox0, n0);
            for (int kb = 0; kb < KB; ++kb, ++g) {
                const uint32_t s = g % S, it = g / S;
                mbar_wait(&empty[s], (it & 1) ^ 1);
                if (elect_one()) {
                    unsigned char* st = smem + s * Cfg::kStageBytes;
                    // K-concatenated 1x1 conv: k-blocks from kb_split on come from the second input (its own pixel stride, no padding)
                    const bool second = kb >= args.kb_split;
                    const CUtensorMap* am = second ? &maps.a2 : &maps.a;
                    const int tap = second ? 0 : kb / cblocks, cb = second ? kb - args.kb_split : kb % cblocks;
                    const int r = tap / args.ksize, ss = tap % args.ksize;
                    const int x = second ? ox0 * args.stride2 : ox0 * args.stride - args.pad + ss;
                    const int y = second ? oy0 * args.stride2 : oy0 * args.stride - args.pad + r;
                    mbar_arrive_expect_tx(&full[s], (uint32_t)Cfg::kStageBytes);
                    if (args.mode & 1) {
                        // 7x7 / stride-2 stem over the zero-haloed NHWC4 input: the 32 floats of a row are the 8 taps x 4 channels of
                        // ONE filter row for one output pixel (overlapping windows, nets.cu launch_f16_stem); a k-block = filter rows
                        // 2 kb and 2 kb + 1 (row 7 carries zero weights)
                        tma_load_4d(st, &maps.a, &full[s], 0, ox0, oy0 * 2 + 2 * kb, b);
                        tma_load_4d(st + 16384, &maps.a, &full[s], 0, ox0, oy0 * 2 + 2 * kb + 1, b);
                    } else {
                        tma_load_4d(st, am, &full[s], cb * kBfBK, x, y, b);
                        tma_load_4d(st + 16384, am, &full[s], cb * kBfBK + 32, x, y, b);
                    }
                    tma_load_2d(st + 32768, &maps.b_hi, &full[s], kb * kBfBK, n0);
                    tma_load_2d(st + 32768 + Cfg::kBBytes, &maps.b_lo, &full[s], kb * kBfBK, n0);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer: converged warp, one elected lane issues (see elect_one)
        uint32_t g = 0, ti = 0;
        for (int id = blockIdx.x; id < total; id += gridDim.x, ++ti) {
            const uint32_t set = SETS == 2 ? (ti & 1) : 0, use = SETS == 2 ? (ti >> 1) : ti;
            mbar_wait(&tmem_empty[set], (use & 1) ^ 1);     // the epilogue drained this accumulator set
            tc_fence_after();
            const uint32_t acc = tmem_base + set * (uint32_t)(NACC * BN);
            for (int kb = 0; kb < KB; ++kb, ++g) {
                const uint32_t s = g % S, it = g / S, slot = g % NSLOT;
                f16_wait(&full[s], it & 1, spin);                       // weight tiles landed
                f16_wait(&a_ready[slot], (g / NSLOT) & 1, spin);        // A_hi / A_lo in the TMEM slot (=> the split warps have read the stage)
                tc_fence_after();
                if (elect_one()) {
                    f16_issue_kblock<BN, NACC>(acc, tmem_base + Cfg::kAcol + slot * 64u, smem_u32(smem + s * Cfg::kStageBytes + 32768), kb);
                    tc_commit(&empty[s]);
                    tc_commit(&a_free[slot]);
                    if (kb == KB - 1) tc_commit(&tmem_full[set]);
                }
                __syncwarp();
            }
        }
    } else if (warp < 4) {
        // idle half of warpgroup 0 (keeps the role groups warpgroup-aligned for setmaxnreg)
    } else if (warp < 12) {
        // ---- split warps: thread = tile row = TMEM lane; warps 4-7 take channels 0..31 of the k-block, warps 8-11 channels 32..63
        f16_regs<kRegsSplit, false>();
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        uint32_t g = 0;
        for (int id = blockIdx.x; id < total; id += gridDim.x) {
            for (int kb = 0; kb < KB; ++kb, ++g) {
                const uint32_t s = g % S, it = g / S, slot = g % NSLOT;
                f16_wait(&full[s], it & 1, spin);
                uint32_t hi[16], lo[16];
                f16_split_row(reinterpret_cast<const float4*>(smem + s * Cfg::kStageBytes + half * 16384 + row * 128), row & 7, hi, lo);
                f16_wait(&a_free[slot], ((g / NSLOT) & 1) ^ 1, spin);   // the MMAs of k-block g - NSLOT released this TMEM slot
                tc_fence_after();
                f16_store_slot(tmem_base + ((uint32_t)(q * 32) << 16) + Cfg::kAcol + slot * 64u + (uint32_t)half * 16u, hi, lo, &a_ready[slot], lane);
            }
        }
    } else {
        f16_regs<kRegsEpilogue, true>();
        f16_epilogue<BN, NACC, SETS>(args, staging, tmem_full, tmem_empty, tmem_base, warp, lane, total, n_tiles);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// 3x3 / stride 1 / pad 1 variant with a HALO tile: the nine taps of a 64-channel slice read shifted windows of ONE 10 x 18
// pixel activation tile instead of nine separately fetched 8 x 16 tiles: the activation bytes an SM takes in per slice drop from
// 9 x 32 KB to 45 KB (layer1/2: 48 -> 21 KB per tap, layer3/4: 64 -> 37 KB per tap).
//   smem   two halo buffers (2 x 23 KB halves each, rows = halo pixels of 32 fp32 channels, SWIZZLE_128B) | weight stages
//          (w_hi | w_lo for one tap of one slice) | epilogue staging
//   warp 0 producer: per slice one halo (two 4-D boxes {32 ch, 18, 10, 1} at (ox0-1, oy0-1): zero fill = padding), then nine
//          weight k-blocks;  warps 2-5 split: tile row (ry, rx) reads halo pixel (ry + r, rx + s) per tap
template <int BN, int NACC, int NSLOT>
__global__ void __launch_bounds__(kBfThreads, 1)
conv_f16_halo_kernel(const __grid_constant__ TcMaps maps, const TcArgs args) {
    using Cfg = F16Cfg<BN, NACC, NSLOT>;
    constexpr int S = Cfg::kStagesB;
    constexpr int SETS = Cfg::kSets;
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    unsigned char* smem = tc_smem_raw;
    unsigned char* bstage = smem + 2 * kHaloBytes;
    float* staging = reinterpret_cast<float*>(bstage + S * Cfg::kBStageBytes);
    uint64_t* bars = (uint64_t*)(bstage + S * Cfg::kBStageBytes + Cfg::kStagingBytes);
    uint64_t* b_full = bars;                   // [S] weight tiles landed
    uint64_t* b_empty = bars + S;              // [S] MMAs finished reading the stage
    uint64_t* h_full = bars + 2 * S;           // [2] halo landed
    uint64_t* h_empty = bars + 2 * S + 2;      // [2] split warps finished the nine taps of the halo
    uint64_t* a_ready = bars + 2 * S + 4;      // [NSLOT]
    uint64_t* a_free = a_ready + NSLOT;        // [NSLOT]
    uint64_t* tmem_full = a_free + NSLOT;      // [2]
    uint64_t* tmem_empty = tmem_full + 2;      // [2]
    uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = args.Cout / BN;
    const int total = args.tiles_x * args.tiles_y * args.B * n_tiles;
    const int cblocks = args.Cin / kBfBK;
    const bool spin = (args.mode & 4) == 0;

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) __trap();
        for (int s = 0; s < S; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&h_full[i], 1);
            mbar_init(&h_empty[i], 8);
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 8);
        }
        for (int i = 0; i < NSLOT; ++i) {
            mbar_init(&a_ready[i], 8);
            mbar_init(&a_free[i], 1);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) f16_regs<kRegsIssue, false>();
    if (warp == 0) {
        // ---- TMA producer (converged warp, one elected lane issues)
        if (elect_one()) {
            tma_prefetch_desc(&maps.a);
            tma_prefetch_desc(&maps.b_hi);
            tma_prefetch_desc(&maps.b_lo);
        }
        uint32_t g = 0, hs = 0;      // weight k-blocks / halo slices issued so far
        for (int id = blockIdx.x; id < total; id += gridDim.x) {
            int b, oy0, ox0, n0;
            f16_tile_coords(args, n_tiles, BN, id, b, oy0, ox0, n0);
            for (int cb = 0; cb < cblocks; ++cb, ++hs) {
                const uint32_t hb = hs & 1;
                mbar_wait(&h_empty[hb], ((hs >> 1) & 1) ^ 1);
                if (elect_one()) {
                    unsigned char* hp = smem + hb * kHaloBytes;
                    mbar_arrive_expect_tx(&h_full[hb], 2u * 32u * 4u * kHaloW * kHaloH);
                    tma_load_4d(hp, &maps.a, &h_full[hb], cb * kBfBK, ox0 - 1, oy0 - 1, b);
                    tma_load_4d(hp + kHaloHalfBytes, &maps.a, &h_full[hb], cb * kBfBK + 32, ox0 - 1, oy0 - 1, b);
                }
                __syncwarp();
                for (int tap = 0; tap < 9; ++tap, ++g) {
                    const uint32_t s = g % S, it = g / S;
                    mbar_wait(&b_empty[s], (it & 1) ^ 1);
                    if (elect_one()) {
                        unsigned char* st = bstage + s * Cfg::kBStageBytes;
                        mbar_arrive_expect_tx(&b_full[s], (uint32_t)Cfg::kBStageBytes);
                        const int kcol = (tap * cblocks + cb) * kBfBK;      // weights are [Cout][tap][Cin]
                        tma_load_2d(st, &maps.b_hi, &b_full[s], kcol, n0);
                        tma_load_2d(st + Cfg::kBBytes, &maps.b_lo, &b_full[s], kcol, n0);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer (converged warp, one elected lane issues)
        uint32_t g = 0, ti = 0;
        const int KB = 9 * cblocks;
        for (int id = blockIdx.x; id < total; id += gridDim.x, ++ti) {
            const uint32_t set = SETS == 2 ? (ti & 1) : 0, use = SETS == 2 ? (ti >> 1) : ti;
            mbar_wait(&tmem_empty[set], (use & 1) ^ 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + set * (uint32_t)(NACC * BN);
            for (int kb = 0; kb < KB; ++kb, ++g) {
                const uint32_t s = g % S, it = g / S, slot = g % NSLOT;
                f16_wait(&b_full[s], it & 1, spin);
                f16_wait(&a_ready[slot], (g / NSLOT) & 1, spin);
                tc_fence_after();
                if (elect_one()) {
                    f16_issue_kblock<BN, NACC>(acc, tmem_base + Cfg::kAcol + slot * 64u, smem_u32(bstage + s * Cfg::kBStageBytes), kb);
                    tc_commit(&b_empty[s]);
                    tc_commit(&a_free[slot]);
                    if (kb == KB - 1) tc_commit(&tmem_full[set]);
                }
                __syncwarp();
            }
        }
    } else if (warp < 4) {
        // idle half of warpgroup 0
    } else if (warp < 12) {
        // ---- split warps: thread = tile row (ry, rx) = TMEM lane, warps 4-7 / 8-11 = channel halves; tap (r, s) reads halo pixel
        // (ry + r, rx + s)
        f16_regs<kRegsSplit, false>();
        const int q = warp & 3;
        const int half = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const int ry = row / kTcTW, rx = row % kTcTW;
        uint32_t g = 0, hs = 0;
        for (int id = blockIdx.x; id < total; id += gridDim.x) {
            for (int cb = 0; cb < cblocks; ++cb, ++hs) {
                const uint32_t hb = hs & 1;
                f16_wait(&h_full[hb], (hs >> 1) & 1, spin);
                const unsigned char* hp = smem + hb * kHaloBytes + half * kHaloHalfBytes;
#pragma unroll 1
                for (int tap = 0; tap < 9; ++tap, ++g) {
                    const uint32_t slot = g % NSLOT;
                    const int p = (ry + tap / 3) * kHaloW + rx + tap % 3;      // halo pixel = 128-byte row of the half
                    uint32_t hi[16], lo[16];
                    f16_split_row(reinterpret_cast<const float4*>(hp + p * 128), p & 7, hi, lo);
                    f16_wait(&a_free[slot], ((g / NSLOT) & 1) ^ 1, spin);
                    tc_fence_after();
                    f16_store_slot(tmem_base + ((uint32_t)(q * 32) << 16) + Cfg::kAcol + slot * 64u + (uint32_t)half * 16u, hi, lo, &a_ready[slot], lane);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&h_empty[hb]);      // every lane's last read of this halo buffer is complete
            }
        }
    } else {
        f16_regs<kRegsEpilogue, true>();
        f16_epilogue<BN, NACC, SETS>(args, staging, tmem_full, tmem_empty, tmem_base, warp, lane, total, n_tiles);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

}  // namespace irn
