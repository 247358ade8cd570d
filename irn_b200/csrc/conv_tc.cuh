// tcgen05 implicit-GEMM convolution for sm_100a: NHWC fp32 activations, 3xTF32 split arithmetic.
//
//   out[b,oy,ox,n] = act( sum_{r,s,c} in[b, oy*st-pad+r, ox*st-pad+s, c] * w[n,(r,s,c)] + bias[n] + residual )
//
// GEMM view per CTA: M = one 8x16 spatial tile of output pixels (128 rows), N = BN output channels,
// K = taps x Cin walked in 32-channel slices.  All kernels are warp-specialised (one CTA per SM):
//   TMA producer  4-D box {32 ch, 16 px, 8 rows, 1 image} of the input per tap (zero fill outside the image = zero
//                 padding; element strides give stride-2 convs), 2-D boxes of the host-split weights (hi / lo planes)
//   split warps   A_hi = round_tf32(a), A_lo = a - A_hi of the activation tile, into shared memory or into TMEM
//   MMA issuer    one thread: tcgen05.mma kind::tf32,  D += A_hi*B_hi + A_hi*B_lo + A_lo*B_hi  (fp32 accumulators in TMEM)
//   epilogue      tcgen05.ld -> shared staging -> coalesced + bias, + residual, ReLU float4 stores
// Four kernels share these pieces (dispatch: nets.cu, run_conv):
//   conv_tc_kernel             one tile per CTA, operands from shared memory (the first version; classifier conv today)
//   conv_tc_persist_kernel     persistent tile loop, two accumulator sets, 8 epilogue warps: K <= 640, 128-channel N tiles
//   conv_tc_persist_ts_kernel  the same with the A operand in TMEM: the 64-channel layers and the stem
//   conv_tc_ts_kernel          one tile per CTA, A operand in TMEM, three accumulators: K > 640
// (conv_tc_ts2_kernel: CTA-pair weight multicast experiment, off by default: no gain.)
// Why 3xTF32: plain TF32 misses the 1e-4 CAM parity bar by 20x; the split keeps ~21 mantissa bits.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "tma.cuh"

namespace irn {

constexpr int kTcTW = 16, kTcTH = 8;          // spatial tile: 128 output pixels
constexpr int kTcBK = 32;                     // fp32 channels per k-block = 128 bytes = one swizzle atom
constexpr int kTcThreads = 192;

struct TcMaps {
    CUtensorMap a;      // input  {Cin, W, H, B}
    CUtensorMap b_hi;   // weights {K, Cout}
    CUtensorMap b_lo;
    CUtensorMap a2;     // conv_f16_kernel only: second input of a K-concatenated 1x1 conv (TcArgs::kb_split); unused otherwise
};

struct TcArgs {
    const float* bias;
    const float* residual;
    float* out;
    int B, Ho, Wo, Cout, Cin, ksize, stride, pad, relu;
    int tiles_x, tiles_y;
    int kb_split = 1 << 30;          // conv_f16_kernel only: k-blocks [kb_split, KB) read input `a2` (pixel stride `stride2`): the projection
    int stride2 = 1;                 // shortcut of a bottleneck fused into conv3's reduction (nets.cu, Block::c3ds)
    const float* oscale = nullptr;   // conv_f16.cuh only: per-output-channel factor undoing the weights' power-of-two pre-scale
    int mode;   // 0: NHWC input, one k-block per (tap, 32-channel slice); 1: stem, NHWC4 zero-haloed input, one k-block per filter row
};

constexpr size_t tc_smem_bytes(int BN, int stages) {
    return 1024 /*align*/ + (size_t)stages * (2 * 16384 + 2 * (size_t)BN * 128) + 256;
}

// one lane of a converged warp (elect.sync): see conv_f16.cuh, elect_one
__device__ __forceinline__ bool tc_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B canonical layout: 8-row groups of 1024 B (SBO), LBO = 1 (unused for swizzled K-major)
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);           // start address
    d |= (uint64_t)1 << 16;                                // leading byte offset (16 B units)
    d |= (uint64_t)(1024 >> 4) << 32;                      // stride byte offset
    d |= (uint64_t)1 << 46;                                // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                                // SWIZZLE_128B
    return d;
}

__device__ __forceinline__ constexpr uint32_t tc_idesc(int M, int N) {
    // c_format F32 (1) @4, a/b format TF32 (2) @7/@10, K-major both, N>>3 @17, M>>4 @24
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// tf32 "hi" part of an fp32 value: round to nearest (ties away, like cvt.rna.tf32.f32) on the 13 dropped mantissa bits.
// Two integer-pipe ops: cvt.rna.tf32 is a quarter-rate conversion and made the split warps the bottleneck of the long-K kernels.
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
// The "lo" part x - hi is exact in fp32 (<= 13 significant bits); the tensor core keeps its top 11 bits (error 2^-21 |x|).

__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

// STAGES: smem pipeline depth.  NACC: 3 = two alternating hi*hi accumulators + one for the cross terms; 2 = one hi*hi + cross
// (short K: few accumulation steps, and 2 x BN TMEM columns let several CTAs share an SM so their fixed latencies overlap).
template <int BN, int STAGES, int NACC>
__global__ void __launch_bounds__(kTcThreads, BN == 64 ? (STAGES == 1 ? 3 : (STAGES == 2 ? 2 : 1)) : 1)
conv_tc_kernel(const __grid_constant__ TcMaps maps, const TcArgs args) {
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    unsigned char* smem = tc_smem_raw;   // dynamic smem base is 1024-aligned by the attribute; checked below
    constexpr int kStageBytes = 2 * 16384 + 2 * BN * 128;
    constexpr int kTcStages = STAGES;
    constexpr int kTmemCols = NACC * BN <= 128 ? 128 : (NACC * BN <= 256 ? 256 : 512);   // power-of-two allocation
    constexpr uint32_t kCrossCol = (NACC - 1) * BN;
    uint64_t* bars = (uint64_t*)(smem + kTcStages * kStageBytes);
    uint64_t* full = bars;                    // [S] TMA landed
    uint64_t* split = bars + kTcStages;       // [S] A_hi / A_lo written
    uint64_t* empty = bars + 2 * kTcStages;   // [S] MMAs finished reading
    uint64_t* acc_full = bars + 3 * kTcStages;
    uint32_t* tmem_slot = (uint32_t*)(bars + 3 * kTcStages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // 1-D grid, N-tile fastest: the CTAs that share an activation tile (and write neighbouring channel segments of the same
    // pixels) run at the same time, so the tile is fetched from HBM once and served to the others from L2
    const int n_tiles = args.Cout / BN;
    const int tile = blockIdx.x / n_tiles;
    const int tx = tile % args.tiles_x;
    const int ty = (tile / args.tiles_x) % args.tiles_y;
    const int b = tile / (args.tiles_x * args.tiles_y);
    const int ox0 = tx * kTcTW, oy0 = ty * kTcTH;
    const int n0 = (blockIdx.x % n_tiles) * BN;
    const int cblocks = args.Cin / kTcBK;
    const int KB = args.mode == 1 ? args.ksize : args.ksize * args.ksize * cblocks;

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) __trap();
        for (int s = 0; s < kTcStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&split[s], 4);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) {   // TMEM allocation: whole warp, BN fp32 accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&maps.a);
            tma_prefetch_desc(&maps.b_hi);
            tma_prefetch_desc(&maps.b_lo);
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % kTcStages, it = kb / kTcStages;
                mbar_wait(&empty[s], (it & 1) ^ 1);
                unsigned char* st = smem + s * kStageBytes;
                const int tap = kb / cblocks, cb = kb % cblocks;
                const int r = tap / args.ksize, ss = tap % args.ksize;
                mbar_arrive_expect_tx(&full[s], 16384u + 2u * BN * 128u);
                if (args.mode == 1)   // stem: row kb of the 7x7 filter; the 8 px x 4 ch window of output ox starts at padded px = 2*ox
                    tma_load_4d(st, &maps.a, &full[s], 0, ox0, oy0 * 2 + kb, b);
                else
                    tma_load_4d(st, &maps.a, &full[s], cb * kTcBK, ox0 * args.stride - args.pad + ss, oy0 * args.stride - args.pad + r, b);
                tma_load_2d(st + 32768, &maps.b_hi, &full[s], kb * kTcBK, n0);
                tma_load_2d(st + 32768 + BN * 128, &maps.b_lo, &full[s], kb * kTcBK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = tc_idesc(128, BN);
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % kTcStages, it = kb / kTcStages;
                mbar_wait(&full[s], it & 1);
                mbar_wait(&split[s], it & 1);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(smem + s * kStageBytes), a_lo = a_hi + 16384;
                const uint32_t b_hi = a_hi + 32768, b_lo = b_hi + BN * 128;
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {   // UMMA_K = 8 tf32 = 32 bytes inside the 128-byte swizzle atom
                    const uint64_t da_hi = tc_smem_desc(a_hi + k4 * 32), da_lo = tc_smem_desc(a_lo + k4 * 32);
                    const uint64_t db_hi = tc_smem_desc(b_hi + k4 * 32), db_lo = tc_smem_desc(b_lo + k4 * 32);
                    // The tensor core's fp32 accumulation truncates: its error grows with the number of accumulation
                    // steps into one accumulator (measured ~7e-6 rel. at K=1152 with a single accumulator).  So the
                    // large hi*hi terms alternate between two accumulators (even / odd k-blocks) and the small
                    // cross terms get a third; the epilogue adds the three in IEEE fp32.
                    tc_mma_tf32(tmem_base + (NACC == 3 ? (uint32_t)((kb & 1) * BN) : 0u), da_hi, db_hi, idesc, (kb >= (NACC == 3 ? 2 : 1) || k4 != 0) ? 1u : 0u);
                    tc_mma_tf32(tmem_base + kCrossCol, da_lo, db_hi, idesc, (kb | k4) != 0);
                    tc_mma_tf32(tmem_base + kCrossCol, da_hi, db_lo, idesc, 1);
                }
                tc_commit(&empty[s]);      // arrives when the MMAs above have finished reading the stage
            }
            tc_commit(acc_full);
        }
    } else {
        // ---- split warps
        const int t = threadIdx.x - 64;   // 0..127
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % kTcStages, it = kb / kTcStages;
            mbar_wait(&full[s], it & 1);
            float4* a = reinterpret_cast<float4*>(smem + s * kStageBytes);
            float4* lo = reinterpret_cast<float4*>(smem + s * kStageBytes + 16384);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 v = a[i * 128 + t];
                float4 h, l;
                h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
                l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
                a[i * 128 + t] = h;
                lo[i * 128 + t] = l;
            }
            fence_proxy_async();   // generic-proxy writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(&split[s]);
        }
        // ---- epilogue.  The direct form (thread = tile row, 128-byte pieces of 32 different rows per instruction) is bound
        // by L1 wavefronts: 32 per load/store instruction instead of 4.  So: phase 1 parks the warp's 32 x BN block in shared
        // memory (the pipeline stages are idle once acc_full fired); phase 2 walks the rows with lanes across the channels --
        // bias / residual loads and output stores are full coalesced row segments.  The residual rows are prefetched into
        // registers 16 row-steps at a time, the first batch before the accumulator is even ready.
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        constexpr int kLd = BN + 4;             // padded row stride (floats): conflict-free float4 rows
        constexpr int kLanesPerRow = BN / 4;    // 32 (BN=128) or 16 (BN=64)
        constexpr int kRowsPerIter = 32 / kLanesPerRow;
        constexpr int kIters = 32 / kRowsPerIter;
        constexpr int kRB = 16;
        const int sub = lane / kLanesPerRow, col = (lane % kLanesPerRow) * 4;
        float* stg = reinterpret_cast<float*>(smem) + (size_t)q * 32 * kLd;
        float4 res[kRB];
        auto row_offset = [&](int it, bool& ok) -> size_t {
            const int row = q * 32 + it * kRowsPerIter + sub;
            const int oy = oy0 + row / kTcTW, ox = ox0 + row % kTcTW;
            ok = oy < args.Ho && ox < args.Wo;
            return (((size_t)b * args.Ho + oy) * args.Wo + ox) * args.Cout + n0 + col;
        };
        auto prefetch = [&](int base) {
#pragma unroll
            for (int i = 0; i < kRB; ++i) {
                bool ok;
                const size_t off = row_offset(base + i, ok);
                res[i] = (ok && args.residual) ? __ldg(reinterpret_cast<const float4*>(args.residual + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        prefetch(0);
        float4 bi = make_float4(0.f, 0.f, 0.f, 0.f);
        if (args.bias) bi = __ldg(reinterpret_cast<const float4*>(args.bias + n0 + col));

        mbar_wait(acc_full, 0);
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < BN / 32; ++cc) {
            uint32_t v[32], u[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cc * 32);
            tc_ld32(taddr, v);
            tc_ld32(taddr + kCrossCol, u);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (NACC == 3 && KB >= 2) {   // the odd-k-block accumulator exists only when there is more than one k-block
                uint32_t t2[32];
                tc_ld32(taddr + (uint32_t)BN, t2);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(t2[j]));
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 o;
                o.x = __uint_as_float(v[j]) + __uint_as_float(u[j]);
                o.y = __uint_as_float(v[j + 1]) + __uint_as_float(u[j + 1]);
                o.z = __uint_as_float(v[j + 2]) + __uint_as_float(u[j + 2]);
                o.w = __uint_as_float(v[j + 3]) + __uint_as_float(u[j + 3]);
                *reinterpret_cast<float4*>(stg + lane * kLd + cc * 32 + j) = o;
            }
        }
        __syncwarp();
#pragma unroll 1
        for (int base = 0; base < kIters; base += kRB) {
            if (base > 0) prefetch(base);
#pragma unroll
            for (int i = 0; i < kRB; ++i) {
                bool ok;
                const size_t off = row_offset(base + i, ok);
                if (ok) {
                    float4 o = *reinterpret_cast<const float4*>(stg + ((base + i) * kRowsPerIter + sub) * kLd + col);
                    o.x += bi.x + res[i].x; o.y += bi.y + res[i].y; o.z += bi.z + res[i].z; o.w += bi.w + res[i].w;
                    if (args.relu) {
                        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                    }
                    *reinterpret_cast<float4*>(args.out + off) = o;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent variant for short reductions (K <= ~640: the 1x1 bottleneck convs and layer1's 3x3).  With 2-20 k-blocks per
// tile the one-tile-per-CTA kernel above is dominated by per-CTA fixed latencies (measured: 9.6 us per 128x64 tile of the
// 64->256 conv, tensor pipe 9 % busy, DRAM 27 %, warps parked on barriers).  Here one CTA per SM walks tiles
// c, c+G, c+2G, ... (N-tile fastest) with every stage decoupled:
//   warp 0      TMA producer, runs ahead through the smem ring across tile boundaries
//   warps 2-5   hi/lo split of the activation k-blocks
//   warp 1      MMA issuer; two TMEM accumulator sets (hi*hi | cross terms each) ping-pong between consecutive tiles
//   warps 6-13  epilogue: TMEM -> shared staging -> coalesced (+bias +residual, ReLU) stores, overlapped with the next
//               tile's loads and MMAs.  Two warps per TMEM lane quarter, each taking half of the tile's channels: with four
//               warps the residual rows had to be fetched in two batches of 16 and the second batch's HBM latency (~1 us per
//               tile, 32 KB in flight per SM) was the critical path of the 1x1 expansion convs; with eight, a warp's whole
//               32-row x BN/2 slice (64 KB per SM) is in flight before the accumulator is even ready.
constexpr int kTcPersistThreads = 448;

template <int BN>
struct TcPersistCfg {
    static constexpr int kStages = BN == 128 ? 2 : 3;
    static constexpr int kStageBytes = 2 * 16384 + 2 * BN * 128;
    static constexpr int kCols = BN / 2;            // channels per epilogue warp
    static constexpr int kLd = kCols + 4;           // padded staging row (floats): conflict-free float4 rows
    static constexpr int kStagingBytes = 2 * 128 * kLd * 4;
    static constexpr size_t kSmem = 1024 + (size_t)kStages * kStageBytes + kStagingBytes + 256;
    static constexpr int kTmemCols = 4 * BN <= 256 ? 256 : 512;   // 2 sets x (main + cross) x BN columns
};

template <int BN>
__global__ void __launch_bounds__(kTcPersistThreads, 1)
conv_tc_persist_kernel(const __grid_constant__ TcMaps maps, const TcArgs args) {
    using Cfg = TcPersistCfg<BN>;
    constexpr int S = Cfg::kStages;
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    unsigned char* smem = tc_smem_raw;
    float* staging = reinterpret_cast<float*>(smem + S * Cfg::kStageBytes);
    uint64_t* bars = (uint64_t*)(smem + S * Cfg::kStageBytes + Cfg::kStagingBytes);
    uint64_t* full = bars;                 // [S]
    uint64_t* split = bars + S;            // [S]
    uint64_t* empty = bars + 2 * S;        // [S]
    uint64_t* tmem_full = bars + 3 * S;    // [2] MMA -> epilogue
    uint64_t* tmem_empty = bars + 3 * S + 2;   // [2] epilogue -> MMA
    uint32_t* tmem_slot = (uint32_t*)(bars + 3 * S + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = args.Cout / BN;
    const int m_tiles = args.tiles_x * args.tiles_y * args.B;
    const int total = m_tiles * n_tiles;
    const int cblocks = args.Cin / kTcBK;
    const int KB = args.mode == 1 ? args.ksize : args.ksize * args.ksize * cblocks;   // mode 1 = stem: one k-block per filter row

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) __trap();
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&split[s], 4);
            mbar_init(&empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tmem_full[i], 1);
            mbar_init(&tmem_empty[i], 8);
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(Cfg::kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto tile_coords = [&](int id, int& b, int& oy0, int& ox0, int& n0) {
        const int m = id / n_tiles;
        n0 = (id % n_tiles) * BN;
        ox0 = (m % args.tiles_x) * kTcTW;
        oy0 = ((m / args.tiles_x) % args.tiles_y) * kTcTH;
        b = m / (args.tiles_x * args.tiles_y);
    };

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&maps.a);
            tma_prefetch_desc(&maps.b_hi);
            tma_prefetch_desc(&maps.b_lo);
            uint32_t g = 0;   // k-blocks issued so far (ring position)
            for (int id = blockIdx.x; id < total; id += gridDim.x) {
                int b, oy0, ox0, n0;
                tile_coords(id, b, oy0, ox0, n0);
                for (int kb = 0; kb < KB; ++kb, ++g) {
                    const uint32_t s = g % S, it = g / S;
                    mbar_wait(&empty[s], (it & 1) ^ 1);
                    unsigned char* st = smem + s * Cfg::kStageBytes;
                    const int tap = kb / cblocks, cb = kb % cblocks;
                    const int r = tap / args.ksize, ss = tap % args.ksize;
                    mbar_arrive_expect_tx(&full[s], 16384u + 2u * BN * 128u);
                    if (args.mode == 1)
                        tma_load_4d(st, &maps.a, &full[s], 0, ox0, oy0 * 2 + kb, b);
                    else
                        tma_load_4d(st, &maps.a, &full[s], cb * kTcBK, ox0 * args.stride - args.pad + ss, oy0 * args.stride - args.pad + r, b);
                    tma_load_2d(st + 32768, &maps.b_hi, &full[s], kb * kTcBK, n0);
                    tma_load_2d(st + 32768 + BN * 128, &maps.b_lo, &full[s], kb * kTcBK, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = tc_idesc(128, BN);
            uint32_t g = 0, ti = 0;
            for (int id = blockIdx.x; id < total; id += gridDim.x, ++ti) {
                const uint32_t set = ti & 1, use = ti >> 1;
                mbar_wait(&tmem_empty[set], (use & 1) ^ 1);     // the epilogue drained this accumulator set
                tc_fence_after();
                const uint32_t acc = tmem_base + set * (2u * BN);
                for (int kb = 0; kb < KB; ++kb, ++g) {
                    const uint32_t s = g % S, it = g / S;
                    mbar_wait(&full[s], it & 1);
                    mbar_wait(&split[s], it & 1);
                    tc_fence_after();
                    const uint32_t a_hi = smem_u32(smem + s * Cfg::kStageBytes), a_lo = a_hi + 16384;
                    const uint32_t b_hi = a_hi + 32768, b_lo = b_hi + BN * 128;
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const uint64_t da_hi = tc_smem_desc(a_hi + k4 * 32), da_lo = tc_smem_desc(a_lo + k4 * 32);
                        const uint64_t db_hi = tc_smem_desc(b_hi + k4 * 32), db_lo = tc_smem_desc(b_lo + k4 * 32);
                        tc_mma_tf32(acc, da_hi, db_hi, idesc, (kb | k4) != 0);
                        tc_mma_tf32(acc + BN, da_lo, db_hi, idesc, (kb | k4) != 0);
                        tc_mma_tf32(acc + BN, da_hi, db_lo, idesc, 1);
                    }
                    tc_commit(&empty[s]);
                }
                tc_commit(&tmem_full[set]);
            }
        }
    } else if (warp < 6) {
        // ---- split warps
        const int t = threadIdx.x - 64;   // 0..127
        uint32_t g = 0;
        for (int id = blockIdx.x; id < total; id += gridDim.x) {
            for (int kb = 0; kb < KB; ++kb, ++g) {
                const uint32_t s = g % S, it = g / S;
                mbar_wait(&full[s], it & 1);
                float4* a = reinterpret_cast<float4*>(smem + s * Cfg::kStageBytes);
                float4* lo = reinterpret_cast<float4*>(smem + s * Cfg::kStageBytes + 16384);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 v = a[i * 128 + t];
                    float4 h, l;
                    h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
                    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
                    a[i * 128 + t] = h;
                    lo[i * 128 + t] = l;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&split[s]);
            }
        }
    } else {
        // ---- epilogue warps 6..13: TMEM lane quarter = warp % 4, channel half = (warp - 6) / 4
        const int q = warp & 3;
        const int hf = (warp - 6) >> 2;
        constexpr int kCols = Cfg::kCols;
        constexpr int kLd = Cfg::kLd;
        constexpr int kLanesPerRow = kCols / 4;          // 16 (BN=128) or 8 (BN=64)
        constexpr int kRowsPerIter = 32 / kLanesPerRow;
        constexpr int kIters = 32 / kRowsPerIter;        // 16 or 8 row-steps cover the warp's 32 rows: one residual batch
        const int sub = lane / kLanesPerRow, col = hf * kCols + (lane % kLanesPerRow) * 4;
        float* stg = staging + (size_t)((hf * 4 + q) * 32) * kLd;
        uint32_t ti = 0;
        for (int id = blockIdx.x; id < total; id += gridDim.x, ++ti) {
            int b, oy0, ox0, n0;
            tile_coords(id, b, oy0, ox0, n0);
            const uint32_t set = ti & 1, use = ti >> 1;
            float4 res[kIters];
            // offsets relative to the tile's first pixel fit 32 bits (8 rows x Wo x Cout); validity of the 16 row-steps as a bit mask
            const size_t tile_base = (((size_t)b * args.Ho + oy0) * args.Wo + ox0) * args.Cout + n0 + col;
            auto rel_offset = [&](int it, bool& ok) -> uint32_t {
                const int row = q * 32 + it * kRowsPerIter + sub;
                const int ry = row / kTcTW, rx = row % kTcTW;
                ok = oy0 + ry < args.Ho && ox0 + rx < args.Wo;
                return (uint32_t)(ry * args.Wo + rx) * (uint32_t)args.Cout;
            };
            const float* res_base = args.residual ? args.residual + tile_base : nullptr;
            float* out_base = args.out + tile_base;
            uint32_t okmask = 0;
#pragma unroll
            for (int i = 0; i < kIters; ++i) {
                bool ok;
                const uint32_t off = rel_offset(i, ok);
                okmask |= (ok ? 1u : 0u) << i;
                res[i] = (ok && res_base) ? __ldg(reinterpret_cast<const float4*>(res_base + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float4 bi = make_float4(0.f, 0.f, 0.f, 0.f);
            if (args.bias) bi = __ldg(reinterpret_cast<const float4*>(args.bias + n0 + col));

            mbar_wait(&tmem_full[set], use & 1);
            tc_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < kCols / 16; ++cc) {
                uint32_t v[16], u[16];
                const uint32_t taddr = tmem_base + set * (2u * BN) + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * kCols + cc * 16);
                tc_ld16(taddr, v);
                tc_ld16(taddr + BN, u);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    float4 o;
                    o.x = __uint_as_float(v[j]) + __uint_as_float(u[j]);
                    o.y = __uint_as_float(v[j + 1]) + __uint_as_float(u[j + 1]);
                    o.z = __uint_as_float(v[j + 2]) + __uint_as_float(u[j + 2]);
                    o.w = __uint_as_float(v[j + 3]) + __uint_as_float(u[j + 3]);
                    *reinterpret_cast<float4*>(stg + lane * kLd + cc * 16 + j) = o;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[set]);   // accumulator set free for the tile after next
#pragma unroll
            for (int i = 0; i < kIters; ++i) {
                if ((okmask >> i) & 1u) {
                    bool ok;
                    const uint32_t off = rel_offset(i, ok);
                    float4 o = *reinterpret_cast<const float4*>(stg + (i * kRowsPerIter + sub) * kLd + (lane % kLanesPerRow) * 4);
                    o.x += bi.x + res[i].x; o.y += bi.y + res[i].y; o.z += bi.z + res[i].z; o.w += bi.w + res[i].w;
                    if (args.relu) {
                        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                    }
                    *reinterpret_cast<float4*>(out_base + off) = o;
                }
            }
            __syncwarp();   // staging rows are rewritten by the next tile's phase 1
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::kTmemCols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// TS variant for the long reductions: the A operand is read from TENSOR MEMORY instead of shared memory.
// Why: in SS mode a 128x128x8 tf32 MMA reads 8 KB of operands from shared memory in its 64-cycle slot = the SM's whole
// 128 B/clk shared-memory bandwidth; the TMA writes (48 KB per k-block) and the hi/lo split (16 KB read + 32 KB written)
// come on top: 192 KB per k-block ~ 1500 cycles against 768 cycles of tensor time (ncu: tensor pipe 58 % active).
// Here the split warps read each activation row once from the (swizzled) TMA tile, and write A_hi / A_lo straight into
// TMEM with tcgen05.st; the MMAs take A from TMEM and only B from shared memory: 112 KB of shared traffic per k-block.
//   TMEM columns: [0,384) three accumulators (hi*hi even/odd k-blocks, cross terms); [384,512) two A slots x (hi 32 | lo 32)
constexpr int kTsStages = 4;
constexpr int kTsStageBytes = 16384 + 2 * 128 * 128;      // A raw | B_hi | B_lo
constexpr size_t kTsSmem = 1024 + (size_t)kTsStages * kTsStageBytes + 256;

__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
          "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
          "r"(v[30]), "r"(v[31])
        : "memory");
}

__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

constexpr int kTsThreads = 320;   // TMA + MMA warps, 8 split warps (two per TMEM lane quarter); the first four also run the epilogue

__global__ void __launch_bounds__(kTsThreads, 1)
conv_tc_ts_kernel(const __grid_constant__ TcMaps maps, const TcArgs args) {
    constexpr int BN = 128;
    constexpr int S = kTsStages;
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    unsigned char* smem = tc_smem_raw;
    uint64_t* bars = (uint64_t*)(smem + S * kTsStageBytes);
    uint64_t* full = bars;                 // [S] TMA landed
    uint64_t* split = bars + S;            // [S] A_hi / A_lo of this k-block are in TMEM
    uint64_t* empty = bars + 2 * S;        // [S] MMAs finished reading the stage's B tiles
    uint64_t* a_free = bars + 3 * S;       // [2] MMAs finished reading TMEM A slot
    uint64_t* acc_full = bars + 3 * S + 2;
    uint32_t* tmem_slot = (uint32_t*)(bars + 3 * S + 3);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = args.Cout / BN;
    const int tile = blockIdx.x / n_tiles;
    const int tx = tile % args.tiles_x;
    const int ty = (tile / args.tiles_x) % args.tiles_y;
    const int b = tile / (args.tiles_x * args.tiles_y);
    const int ox0 = tx * kTcTW, oy0 = ty * kTcTH;
    const int n0 = (blockIdx.x % n_tiles) * BN;
    const int cblocks = args.Cin / kTcBK;
    const int KB = args.ksize * args.ksize * cblocks;

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) __trap();
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&split[s], 8);
            mbar_init(&empty[s], 1);
        }
        mbar_init(&a_free[0], 1);
        mbar_init(&a_free[1], 1);
        mbar_init(acc_full, 1);
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
    constexpr uint32_t kAcol = 384;

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&maps.a);
            tma_prefetch_desc(&maps.b_hi);
            tma_prefetch_desc(&maps.b_lo);
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % S, it = kb / S;
                mbar_wait(&empty[s], (it & 1) ^ 1);
                unsigned char* st = smem + s * kTsStageBytes;
                const int tap = kb / cblocks, cb = kb % cblocks;
                const int r = tap / args.ksize, ss = tap % args.ksize;
                mbar_arrive_expect_tx(&full[s], 16384u + 2u * BN * 128u);
                tma_load_4d(st, &maps.a, &full[s], cb * kTcBK, ox0 * args.stride - args.pad + ss, oy0 * args.stride - args.pad + r, b);
                tma_load_2d(st + 16384, &maps.b_hi, &full[s], kb * kTcBK, n0);
                tma_load_2d(st + 16384 + BN * 128, &maps.b_lo, &full[s], kb * kTcBK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = tc_idesc(128, BN);
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % S, it = kb / S;
                mbar_wait(&full[s], it & 1);
                mbar_wait(&split[s], it & 1);
                tc_fence_after();
                const uint32_t b_hi = smem_u32(smem + s * kTsStageBytes + 16384), b_lo = b_hi + BN * 128;
                const uint32_t a_t = tmem_base + kAcol + (uint32_t)(kb & 1) * 64u;
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const uint64_t db_hi = tc_smem_desc(b_hi + k4 * 32), db_lo = tc_smem_desc(b_lo + k4 * 32);
                    const uint32_t ta_hi = a_t + k4 * 8, ta_lo = a_t + 32 + k4 * 8;
                    tc_mma_tf32_ts(tmem_base + (uint32_t)((kb & 1) * BN), ta_hi, db_hi, idesc, (kb >= 2 || k4 != 0) ? 1u : 0u);
                    tc_mma_tf32_ts(tmem_base + 2u * BN, ta_lo, db_hi, idesc, (kb | k4) != 0);
                    tc_mma_tf32_ts(tmem_base + 2u * BN, ta_hi, db_lo, idesc, 1);
                }
                tc_commit(&empty[s]);
                tc_commit(&a_free[kb & 1]);
            }
            tc_commit(acc_full);
        }
    } else {
        // ---- split warps: thread = tile row (TMEM lane); warps 2-5 take the first 16 channels of the row's 128-byte slice of the
        // swizzled TMA tile, warps 6-9 the other 16 (two warps per scheduler keep the LDS -> round -> STTM chain busy)
        const int q = warp & 3;
        const int half = warp >= 6 ? 1 : 0;
        const int row = q * 32 + lane;
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % S, it = kb / S;
            const int slot = kb & 1;
            mbar_wait(&full[s], it & 1);
            const float4* arow = reinterpret_cast<const float4*>(smem + s * kTsStageBytes + row * 128);
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 v = arow[(half * 4 + c) ^ (row & 7)];      // SWIZZLE_128B: 16-byte chunk c of row r lives at chunk c ^ (r % 8)
                const float h0 = tf32_hi(v.x), h1 = tf32_hi(v.y), h2 = tf32_hi(v.z), h3 = tf32_hi(v.w);
                hi[4 * c] = __float_as_uint(h0); hi[4 * c + 1] = __float_as_uint(h1); hi[4 * c + 2] = __float_as_uint(h2); hi[4 * c + 3] = __float_as_uint(h3);
                lo[4 * c] = __float_as_uint(v.x - h0); lo[4 * c + 1] = __float_as_uint(v.y - h1);
                lo[4 * c + 2] = __float_as_uint(v.z - h2); lo[4 * c + 3] = __float_as_uint(v.w - h3);
            }
            mbar_wait(&a_free[slot], ((kb >> 1) & 1) ^ 1);    // the MMAs of k-block kb-2 released this TMEM slot
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + kAcol + (uint32_t)slot * 64u + (uint32_t)half * 16u;
            tc_st16(taddr, hi);
            tc_st16(taddr + 32, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&split[s]);
        }
        {
        // ---- epilogue on all eight split warps: TMEM lane quarter q, channel half `half` (64 of the 128 channels each), so a
        // warp's whole 32-row slice of the residual is in flight in one batch before the accumulator is ready
        constexpr int kCols = BN / 2;
        constexpr int kLd = kCols + 4;
        constexpr int kIters = 16;                       // 2 rows per step (16 lanes x float4 per row)
        const int sub = lane >> 4, col = half * kCols + (lane & 15) * 4;
        float* stg = reinterpret_cast<float*>(smem) + (size_t)((half * 4 + q) * 32) * kLd;
        const size_t tile_base = (((size_t)b * args.Ho + oy0) * args.Wo + ox0) * args.Cout + n0 + col;
        auto rel_offset = [&](int it, bool& ok) -> uint32_t {
            const int rr = q * 32 + it * 2 + sub;
            const int ry = rr / kTcTW, rx = rr % kTcTW;
            ok = oy0 + ry < args.Ho && ox0 + rx < args.Wo;
            return (uint32_t)(ry * args.Wo + rx) * (uint32_t)args.Cout;
        };
        const float* res_base = args.residual ? args.residual + tile_base : nullptr;
        float* out_base = args.out + tile_base;
        float4 res[kIters];
        uint32_t okmask = 0;
#pragma unroll
        for (int i = 0; i < kIters; ++i) {
            bool ok;
            const uint32_t off = rel_offset(i, ok);
            okmask |= (ok ? 1u : 0u) << i;
            res[i] = (ok && res_base) ? __ldg(reinterpret_cast<const float4*>(res_base + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float4 bi = make_float4(0.f, 0.f, 0.f, 0.f);
        if (args.bias) bi = __ldg(reinterpret_cast<const float4*>(args.bias + n0 + col));
        mbar_wait(acc_full, 0);
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < kCols / 16; ++cc) {
            uint32_t v[16], u[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * kCols + cc * 16);
            tc_ld16(taddr, v);
            tc_ld16(taddr + 2u * BN, u);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (KB >= 2) {
                uint32_t t2[16];
                tc_ld16(taddr + (uint32_t)BN, t2);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(t2[j]));
            }
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                float4 o;
                o.x = __uint_as_float(v[j]) + __uint_as_float(u[j]);
                o.y = __uint_as_float(v[j + 1]) + __uint_as_float(u[j + 1]);
                o.z = __uint_as_float(v[j + 2]) + __uint_as_float(u[j + 2]);
                o.w = __uint_as_float(v[j + 3]) + __uint_as_float(u[j + 3]);
                *reinterpret_cast<float4*>(stg + lane * kLd + cc * 16 + j) = o;
            }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < kIters; ++i) {
            if ((okmask >> i) & 1u) {
                bool ok;
                const uint32_t off = rel_offset(i, ok);
                float4 o = *reinterpret_cast<const float4*>(stg + (i * 2 + sub) * kLd + (lane & 15) * 4);
                o.x += bi.x + res[i].x; o.y += bi.y + res[i].y; o.z += bi.z + res[i].z; o.w += bi.w + res[i].w;
                if (args.relu) {
                    o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                }
                *reinterpret_cast<float4*>(out_base + off) = o;
            }
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent kernel with the A operand in tensor memory, for the 64-channel layers with short reductions (layer1's 3x3 and
// 256->64, the stem): the persistent tile loop / double-buffered accumulators / 8-warp epilogue of conv_tc_persist_kernel with
// the split of conv_tc_ts_kernel.  Why: with N = 64 an SS-mode MMA reads 6 KB of operands from shared memory per 32 cycles of
// tensor time (192 B/clk against the 128 B/clk the SM has) and the hi/lo split adds 48 KB per k-block: layer1's 3x3 ran at
// 1255 cycles per k-block = the shared-memory bound (152 KB).  With A in TMEM a k-block moves 72 KB through shared memory (TMA
// 32 + split reads 16 + B operand reads 24), and the smaller stages (32 KB) allow a five-deep ring = 80 KB of activations in
// flight per SM for the memory-bound 256->64 reductions.
//   TMEM columns: [0,256) two accumulator sets x (main 64 | cross 64); [256,384) two A slots x (hi 32 | lo 32)
//   warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 split (thread = tile row), warps 6-13 epilogue
constexpr int kPtsBN = 64;
constexpr int kPtsStages = 5;
constexpr int kPtsStageBytes = 16384 + 2 * kPtsBN * 128;          // A raw | B_hi | B_lo
constexpr int kPtsLd = kPtsBN / 2 + 4;
constexpr int kPtsStagingBytes = 2 * 128 * kPtsLd * 4;
constexpr size_t kPtsSmem = 1024 + (size_t)kPtsStages * kPtsStageBytes + kPtsStagingBytes + 256;

__global__ void __launch_bounds__(kTcPersistThreads, 1)
conv_tc_persist_ts_kernel(const __grid_constant__ TcMaps maps, const TcArgs args) {
    constexpr int BN = kPtsBN;
    constexpr int S = kPtsStages;
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    unsigned char* smem = tc_smem_raw;
    float* staging = reinterpret_cast<float*>(smem + S * kPtsStageBytes);
    uint64_t* bars = (uint64_t*)(smem + S * kPtsStageBytes + kPtsStagingBytes);
    uint64_t* full = bars;                     // [S] TMA landed
    uint64_t* split = bars + S;                // [S] A_hi / A_lo of the k-block are in TMEM
    uint64_t* empty = bars + 2 * S;            // [S] MMAs finished reading the stage
    uint64_t* a_free = bars + 3 * S;           // [2] MMAs finished reading TMEM A slot
    uint64_t* tmem_full = bars + 3 * S + 2;    // [2] MMA -> epilogue
    uint64_t* tmem_empty = bars + 3 * S + 4;   // [2] epilogue -> MMA
    uint32_t* tmem_slot = (uint32_t*)(bars + 3 * S + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = args.Cout / BN;
    const int m_tiles = args.tiles_x * args.tiles_y * args.B;
    const int total = m_tiles * n_tiles;
    const int cblocks = args.Cin / kTcBK;
    const int KB = args.mode == 1 ? args.ksize : args.ksize * args.ksize * cblocks;

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) __trap();
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&split[s], 4);
            mbar_init(&empty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_free[i], 1);
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
    constexpr uint32_t kAcol = 256;

    auto tile_coords = [&](int id, int& b, int& oy0, int& ox0, int& n0) {
        const int m = id / n_tiles;
        n0 = (id % n_tiles) * BN;
        ox0 = (m % args.tiles_x) * kTcTW;
        oy0 = ((m / args.tiles_x) % args.tiles_y) * kTcTH;
        b = m / (args.tiles_x * args.tiles_y);
    };

    // Producer and MMA loops run on a CONVERGED warp and issue under elect.sync: under `if (lane == 0)` the compiler wraps every
    // tcgen05.mma / tcgen05.commit / TMA instruction in its own ELECT + BRA.U.ANY loop (~100 cycles each; conv_f16.cuh, elect_one).
    if (warp == 0) {
        if (tc_elect_one()) {
            tma_prefetch_desc(&maps.a);
            tma_prefetch_desc(&maps.b_hi);
            tma_prefetch_desc(&maps.b_lo);
        }
        uint32_t g = 0;
        for (int id = blockIdx.x; id < total; id += gridDim.x) {
            int b, oy0, ox0, n0;
            tile_coords(id, b, oy0, ox0, n0);
            for (int kb = 0; kb < KB; ++kb, ++g) {
                const uint32_t s = g % S, it = g / S;
                mbar_wait(&empty[s], (it & 1) ^ 1);
                if (tc_elect_one()) {
                    unsigned char* st = smem + s * kPtsStageBytes;
                    const int tap = kb / cblocks, cb = kb % cblocks;
                    const int r = tap / args.ksize, ss = tap % args.ksize;
                    mbar_arrive_expect_tx(&full[s], 16384u + 2u * BN * 128u);
                    if (args.mode == 1)
                        tma_load_4d(st, &maps.a, &full[s], 0, ox0, oy0 * 2 + kb, b);
                    else
                        tma_load_4d(st, &maps.a, &full[s], cb * kTcBK, ox0 * args.stride - args.pad + ss, oy0 * args.stride - args.pad + r, b);
                    tma_load_2d(st + 16384, &maps.b_hi, &full[s], kb * kTcBK, n0);
                    tma_load_2d(st + 16384 + BN * 128, &maps.b_lo, &full[s], kb * kTcBK, n0);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = tc_idesc(128, BN);
        uint32_t g = 0, ti = 0;
        for (int id = blockIdx.x; id < total; id += gridDim.x, ++ti) {
            const uint32_t set = ti & 1, use = ti >> 1;
            mbar_wait(&tmem_empty[set], (use & 1) ^ 1);
            tc_fence_after();
            const uint32_t acc = tmem_base + set * (2u * BN);
            for (int kb = 0; kb < KB; ++kb, ++g) {
                const uint32_t s = g % S, it = g / S, slot = g & 1;
                mbar_wait(&full[s], it & 1);       // B tiles landed
                mbar_wait(&split[s], it & 1);      // A_hi / A_lo in TMEM slot
                tc_fence_after();
                if (tc_elect_one()) {
                    const uint32_t b_hi = smem_u32(smem + s * kPtsStageBytes + 16384), b_lo = b_hi + BN * 128;
                    const uint32_t a_t = tmem_base + kAcol + slot * 64u;
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        const uint64_t db_hi = tc_smem_desc(b_hi + k4 * 32), db_lo = tc_smem_desc(b_lo + k4 * 32);
                        const uint32_t ta_hi = a_t + k4 * 8, ta_lo = a_t + 32 + k4 * 8;
                        tc_mma_tf32_ts(acc, ta_hi, db_hi, idesc, (kb | k4) != 0);
                        tc_mma_tf32_ts(acc + BN, ta_lo, db_hi, idesc, (kb | k4) != 0);
                        tc_mma_tf32_ts(acc + BN, ta_hi, db_lo, idesc, 1);
                    }
                    tc_commit(&empty[s]);
                    tc_commit(&a_free[slot]);
                    if (kb == KB - 1) tc_commit(&tmem_full[set]);
                }
                __syncwarp();
            }
        }
    } else if (warp < 6) {
        // ---- split warps: thread = tile row = TMEM lane; the row's 128-byte slice of the swizzled TMA tile -> hi / lo -> TMEM
        const int q = warp & 3;
        const int row = q * 32 + lane;
        uint32_t g = 0;
        for (int id = blockIdx.x; id < total; id += gridDim.x) {
            for (int kb = 0; kb < KB; ++kb, ++g) {
                const uint32_t s = g % S, it = g / S, slot = g & 1;
                mbar_wait(&full[s], it & 1);
                const float4* arow = reinterpret_cast<const float4*>(smem + s * kPtsStageBytes + row * 128);
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 v = arow[c ^ (row & 7)];      // SWIZZLE_128B: 16-byte chunk c of row r lives at chunk c ^ (r % 8)
                    const float h0 = tf32_hi(v.x), h1 = tf32_hi(v.y), h2 = tf32_hi(v.z), h3 = tf32_hi(v.w);
                    hi[4 * c] = __float_as_uint(h0); hi[4 * c + 1] = __float_as_uint(h1); hi[4 * c + 2] = __float_as_uint(h2); hi[4 * c + 3] = __float_as_uint(h3);
                    lo[4 * c] = __float_as_uint(v.x - h0); lo[4 * c + 1] = __float_as_uint(v.y - h1);
                    lo[4 * c + 2] = __float_as_uint(v.z - h2); lo[4 * c + 3] = __float_as_uint(v.w - h3);
                }
                mbar_wait(&a_free[slot], ((g >> 1) & 1) ^ 1);   // the MMAs of k-block g-2 released this TMEM slot
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + kAcol + slot * 64u;
                tc_st32(taddr, hi);
                tc_st32(taddr + 32, lo);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&split[s]);
            }
        }
    } else {
        // ---- epilogue warps 6..13 (as in conv_tc_persist_kernel): TMEM lane quarter = warp % 4, channel half = (warp - 6) / 4
        const int q = warp & 3;
        const int hf = (warp - 6) >> 2;
        constexpr int kCols = BN / 2;
        constexpr int kLd = kPtsLd;
        constexpr int kLanesPerRow = kCols / 4;          // 8
        constexpr int kRowsPerIter = 32 / kLanesPerRow;  // 4
        constexpr int kIters = 32 / kRowsPerIter;        // 8
        const int sub = lane / kLanesPerRow, col = hf * kCols + (lane % kLanesPerRow) * 4;
        float* stg = staging + (size_t)((hf * 4 + q) * 32) * kLd;
        uint32_t ti = 0;
        for (int id = blockIdx.x; id < total; id += gridDim.x, ++ti) {
            int b, oy0, ox0, n0;
            tile_coords(id, b, oy0, ox0, n0);
            const uint32_t set = ti & 1, use = ti >> 1;
            float4 res[kIters];
            const size_t tile_base = (((size_t)b * args.Ho + oy0) * args.Wo + ox0) * args.Cout + n0 + col;
            auto rel_offset = [&](int it, bool& ok) -> uint32_t {
                const int rr = q * 32 + it * kRowsPerIter + sub;
                const int ry = rr / kTcTW, rx = rr % kTcTW;
                ok = oy0 + ry < args.Ho && ox0 + rx < args.Wo;
                return (uint32_t)(ry * args.Wo + rx) * (uint32_t)args.Cout;
            };
            const float* res_base = args.residual ? args.residual + tile_base : nullptr;
            float* out_base = args.out + tile_base;
            uint32_t okmask = 0;
#pragma unroll
            for (int i = 0; i < kIters; ++i) {
                bool ok;
                const uint32_t off = rel_offset(i, ok);
                okmask |= (ok ? 1u : 0u) << i;
                res[i] = (ok && res_base) ? __ldg(reinterpret_cast<const float4*>(res_base + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float4 bi = make_float4(0.f, 0.f, 0.f, 0.f);
            if (args.bias) bi = __ldg(reinterpret_cast<const float4*>(args.bias + n0 + col));

            mbar_wait(&tmem_full[set], use & 1);
            tc_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < kCols / 16; ++cc) {
                uint32_t v[16], u[16];
                const uint32_t taddr = tmem_base + set * (2u * BN) + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * kCols + cc * 16);
                tc_ld16(taddr, v);
                tc_ld16(taddr + BN, u);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    float4 o;
                    o.x = __uint_as_float(v[j]) + __uint_as_float(u[j]);
                    o.y = __uint_as_float(v[j + 1]) + __uint_as_float(u[j + 1]);
                    o.z = __uint_as_float(v[j + 2]) + __uint_as_float(u[j + 2]);
                    o.w = __uint_as_float(v[j + 3]) + __uint_as_float(u[j + 3]);
                    *reinterpret_cast<float4*>(stg + lane * kLd + cc * 16 + j) = o;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[set]);
#pragma unroll
            for (int i = 0; i < kIters; ++i) {
                if ((okmask >> i) & 1u) {
                    bool ok;
                    const uint32_t off = rel_offset(i, ok);
                    float4 o = *reinterpret_cast<const float4*>(stg + (i * kRowsPerIter + sub) * kLd + (lane % kLanesPerRow) * 4);
                    o.x += bi.x + res[i].x; o.y += bi.y + res[i].y; o.z += bi.z + res[i].z; o.w += bi.w + res[i].w;
                    if (args.relu) {
                        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                    }
                    *reinterpret_cast<float4*>(out_base + off) = o;
                }
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}

#ifdef IRN_EXPERIMENTAL   // CTA-pair weight multicast variant of the long-K kernel: measured no gain (profiles/r01_conv_micro_final.txt)
// ---- 2-CTA cluster helpers (weight tile multicast)
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// Same as conv_tc_ts_kernel, in CTA pairs: the L2 -> SM operand feed (48 KB per k-block per SM, 6.2 KB/clk chip-wide = the L2
// slice limit) was the binding constraint of the long-K layers once the A operand moved to TMEM; sharing the weight tile
// between two M tiles cuts it to 32 KB.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTsThreads, 1)
conv_tc_ts2_kernel(const __grid_constant__ TcMaps maps, const TcArgs args) {
    constexpr int BN = 128;
    constexpr int S = kTsStages;
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    unsigned char* smem = tc_smem_raw;
    uint64_t* bars = (uint64_t*)(smem + S * kTsStageBytes);
    uint64_t* full = bars;                 // [S] TMA landed
    uint64_t* split = bars + S;            // [S] A_hi / A_lo of this k-block are in TMEM
    uint64_t* empty = bars + 2 * S;        // [S] MMAs finished reading the stage's B tiles
    uint64_t* a_free = bars + 3 * S;       // [2] MMAs finished reading TMEM A slot
    uint64_t* acc_full = bars + 3 * S + 2;
    uint32_t* tmem_slot = (uint32_t*)(bars + 3 * S + 3);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // CTA pair = two consecutive M tiles of the same N tile: each CTA fetches half of the weight tile and multicasts it to both
    const uint32_t rank = cluster_rank();
    const int n_tiles = args.Cout / BN;
    const int pair = blockIdx.x >> 1;
    const int m_tiles = args.tiles_x * args.tiles_y * args.B;
    const int tile = (pair / n_tiles) * 2 + (int)rank;
    const bool ghost = tile >= m_tiles;            // odd tile count: the partner of the last tile computes on zero-filled input, stores nothing
    const int tx = tile % args.tiles_x;
    const int ty = (tile / args.tiles_x) % args.tiles_y;
    const int b = tile / (args.tiles_x * args.tiles_y);   // == args.B for the ghost: TMA zero-fills
    const int ox0 = tx * kTcTW, oy0 = ty * kTcTH;
    const int n0 = (pair % n_tiles) * BN;
    const int cblocks = args.Cin / kTcBK;
    const int KB = args.ksize * args.ksize * cblocks;

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) __trap();
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&split[s], 8);
            mbar_init(&empty[s], 2);        // both CTAs of the pair must have finished with stage s before it is refilled
        }
        mbar_init(&a_free[0], 1);
        mbar_init(&a_free[1], 1);
        mbar_init(acc_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                            // the peer's barriers exist before anything is multicast at them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t kAcol = 384;

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&maps.a);
            tma_prefetch_desc(&maps.b_hi);
            tma_prefetch_desc(&maps.b_lo);
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % S, it = kb / S;
                mbar_wait(&empty[s], (it & 1) ^ 1);
                unsigned char* st = smem + s * kTsStageBytes;
                const int tap = kb / cblocks, cb = kb % cblocks;
                const int r = tap / args.ksize, ss = tap % args.ksize;
                mbar_arrive_expect_tx(&full[s], 16384u + 2u * BN * 128u);
                tma_load_4d(st, &maps.a, &full[s], cb * kTcBK, ox0 * args.stride - args.pad + ss, oy0 * args.stride - args.pad + r, b);
                // this CTA's 64 rows of the weight tile go to the same place in BOTH CTAs (and signal both full barriers)
                tma_load_2d_mc(st + 16384 + rank * 8192, &maps.b_hi, &full[s], kb * kTcBK, n0 + (int)rank * 64, (uint16_t)3);
                tma_load_2d_mc(st + 16384 + BN * 128 + rank * 8192, &maps.b_lo, &full[s], kb * kTcBK, n0 + (int)rank * 64, (uint16_t)3);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = tc_idesc(128, BN);
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % S, it = kb / S;
                mbar_wait(&full[s], it & 1);
                mbar_wait(&split[s], it & 1);
                tc_fence_after();
                const uint32_t b_hi = smem_u32(smem + s * kTsStageBytes + 16384), b_lo = b_hi + BN * 128;
                const uint32_t a_t = tmem_base + kAcol + (uint32_t)(kb & 1) * 64u;
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const uint64_t db_hi = tc_smem_desc(b_hi + k4 * 32), db_lo = tc_smem_desc(b_lo + k4 * 32);
                    const uint32_t ta_hi = a_t + k4 * 8, ta_lo = a_t + 32 + k4 * 8;
                    tc_mma_tf32_ts(tmem_base + (uint32_t)((kb & 1) * BN), ta_hi, db_hi, idesc, (kb >= 2 || k4 != 0) ? 1u : 0u);
                    tc_mma_tf32_ts(tmem_base + 2u * BN, ta_lo, db_hi, idesc, (kb | k4) != 0);
                    tc_mma_tf32_ts(tmem_base + 2u * BN, ta_hi, db_lo, idesc, 1);
                }
                tc_commit_mc(&empty[s], (uint16_t)3);   // frees stage s in both CTAs
                tc_commit(&a_free[kb & 1]);
            }
            tc_commit(acc_full);
        }
    } else {
        // ---- split warps: thread = tile row (TMEM lane); warps 2-5 take the first 16 channels of the row's 128-byte slice of the
        // swizzled TMA tile, warps 6-9 the other 16 (two warps per scheduler keep the LDS -> round -> STTM chain busy)
        const int q = warp & 3;
        const int half = warp >= 6 ? 1 : 0;
        const int row = q * 32 + lane;
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % S, it = kb / S;
            const int slot = kb & 1;
            mbar_wait(&full[s], it & 1);
            const float4* arow = reinterpret_cast<const float4*>(smem + s * kTsStageBytes + row * 128);
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 v = arow[(half * 4 + c) ^ (row & 7)];      // SWIZZLE_128B: 16-byte chunk c of row r lives at chunk c ^ (r % 8)
                const float h0 = tf32_hi(v.x), h1 = tf32_hi(v.y), h2 = tf32_hi(v.z), h3 = tf32_hi(v.w);
                hi[4 * c] = __float_as_uint(h0); hi[4 * c + 1] = __float_as_uint(h1); hi[4 * c + 2] = __float_as_uint(h2); hi[4 * c + 3] = __float_as_uint(h3);
                lo[4 * c] = __float_as_uint(v.x - h0); lo[4 * c + 1] = __float_as_uint(v.y - h1);
                lo[4 * c + 2] = __float_as_uint(v.z - h2); lo[4 * c + 3] = __float_as_uint(v.w - h3);
            }
            mbar_wait(&a_free[slot], ((kb >> 1) & 1) ^ 1);    // the MMAs of k-block kb-2 released this TMEM slot
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + kAcol + (uint32_t)slot * 64u + (uint32_t)half * 16u;
            tc_st16(taddr, hi);
            tc_st16(taddr + 32, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&split[s]);
        }
        if (warp < 6) {
        // ---- epilogue on warps 2-5 (same as conv_tc_kernel: TMEM -> shared staging -> coalesced stores, residual rows prefetched)
        constexpr int kLd = BN + 4;
        constexpr int kRB = 16;
        const int col = lane * 4;
        float* stg = reinterpret_cast<float*>(smem) + (size_t)q * 32 * kLd;
        float4 res[kRB];
        auto row_offset = [&](int it, bool& ok) -> size_t {
            const int rr = q * 32 + it;
            const int oy = oy0 + rr / kTcTW, ox = ox0 + rr % kTcTW;
            ok = !ghost && oy < args.Ho && ox < args.Wo;
            return (((size_t)b * args.Ho + oy) * args.Wo + ox) * args.Cout + n0 + col;
        };
        auto prefetch = [&](int base) {
#pragma unroll
            for (int i = 0; i < kRB; ++i) {
                bool ok;
                const size_t off = row_offset(base + i, ok);
                res[i] = (ok && args.residual) ? __ldg(reinterpret_cast<const float4*>(args.residual + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        prefetch(0);
        float4 bi = make_float4(0.f, 0.f, 0.f, 0.f);
        if (args.bias) bi = __ldg(reinterpret_cast<const float4*>(args.bias + n0 + col));
        mbar_wait(acc_full, 0);
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < BN / 32; ++cc) {
            uint32_t v[32], u[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cc * 32);
            tc_ld32(taddr, v);
            tc_ld32(taddr + 2u * BN, u);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (KB >= 2) {
                uint32_t t2[32];
                tc_ld32(taddr + (uint32_t)BN, t2);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(t2[j]));
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 o;
                o.x = __uint_as_float(v[j]) + __uint_as_float(u[j]);
                o.y = __uint_as_float(v[j + 1]) + __uint_as_float(u[j + 1]);
                o.z = __uint_as_float(v[j + 2]) + __uint_as_float(u[j + 2]);
                o.w = __uint_as_float(v[j + 3]) + __uint_as_float(u[j + 3]);
                *reinterpret_cast<float4*>(stg + lane * kLd + cc * 32 + j) = o;
            }
        }
        __syncwarp();
#pragma unroll 1
        for (int base = 0; base < 32; base += kRB) {
            if (base > 0) prefetch(base);
#pragma unroll
            for (int i = 0; i < kRB; ++i) {
                bool ok;
                const size_t off = row_offset(base + i, ok);
                if (ok) {
                    float4 o = *reinterpret_cast<const float4*>(stg + (base + i) * kLd + col);
                    o.x += bi.x + res[i].x; o.y += bi.y + res[i].y; o.z += bi.z + res[i].z; o.w += bi.w + res[i].w;
                    if (args.relu) {
                        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                    }
                    *reinterpret_cast<float4*>(args.out + off) = o;
                }
            }
        }
        }   // warp < 6
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                            // nobody exits while the peer may still write into its shared memory / barriers
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
}


#endif  // IRN_EXPERIMENTAL

}  // namespace irn
