// P1-P4: instance path -- centroid refinement on the displacement field, connected components, centroid
// clustering, per-instance seeds and per-segment statistics.
// Reference: step/make_ins_seg_labels.py:18-105 (+ misc/imutils.py:182-190, misc/pyutils.py:86-101).
#include "common.h"

namespace irn {

// ---------------------------------------------------------------- P1 centroids
// find_centroids_with_refinement (step/make_ins_seg_labels.py:18-56).  numpy evaluates the bilinear update in
// float64 (float32 centroid - int32 floor promotes), four products summed left to right, no FMA; the in-place
// `+=` rounds the float64 sum back to float32; np.clip; final np.round (half to even).  Mirrored op for op.
__global__ void centroids_kernel(const float* __restrict__ dp, int32_t* __restrict__ out, int h, int w, int iterations) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= h * w) return;
    const float* d0 = dp;
    const float* d1 = dp + (size_t)h * w;
    float cy = (float)(p / w), cx = (float)(p % w);
    const float ymax = (float)(h - 1), xmax = (float)(w - 1);
    for (int it = 0; it < iterations; ++it) {
        const int uy = (int)ceilf(cy), ly = (int)floorf(cy);
        const int ux = (int)ceilf(cx), lx = (int)floorf(cx);
        const double yc = __dsub_rn((double)cy, (double)ly), xc = __dsub_rn((double)cx, (double)lx);
        const double yn = __dsub_rn(1.0, yc), xn = __dsub_rn(1.0, xc);
        const int i_uu = uy * w + ux, i_lu = ly * w + ux, i_ul = uy * w + lx, i_ll = ly * w + lx;
        double sy = __dmul_rn(__dmul_rn((double)d0[i_uu], yc), xc);
        sy = __dadd_rn(sy, __dmul_rn(__dmul_rn((double)d0[i_lu], yn), xc));
        sy = __dadd_rn(sy, __dmul_rn(__dmul_rn((double)d0[i_ul], yc), xn));
        sy = __dadd_rn(sy, __dmul_rn(__dmul_rn((double)d0[i_ll], yn), xn));
        double sx = __dmul_rn(__dmul_rn((double)d1[i_uu], yc), xc);
        sx = __dadd_rn(sx, __dmul_rn(__dmul_rn((double)d1[i_lu], yn), xc));
        sx = __dadd_rn(sx, __dmul_rn(__dmul_rn((double)d1[i_ul], yc), xn));
        sx = __dadd_rn(sx, __dmul_rn(__dmul_rn((double)d1[i_ll], yn), xn));
        cy = (float)__dadd_rn((double)cy, sy);
        cx = (float)__dadd_rn((double)cx, sx);
        cy = fminf(fmaxf(cy, 0.f), ymax);
        cx = fminf(fmaxf(cx, 0.f), xmax);
    }
    out[p] = (int)rintf(cy);
    out[(size_t)h * w + p] = (int)rintf(cx);
}

// ---------------------------------------------------------------- connected components (4-connectivity)
// Union-find with atomicMin links (larger root -> smaller root): the root of a component is its minimum linear
// index, i.e. its first pixel in raster order -- the numbering order of skimage.measure.label / scipy.ndimage.label.
__device__ __forceinline__ int uf_find(const int* L, int x) {
    int y = L[x];
    while (y != x) {
        x = y;
        y = L[x];
    }
    return x;
}

__device__ __forceinline__ void uf_union(int* L, int a, int b) {
    while (true) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) return;
        if (a < b) {
            const int t = a;
            a = b;
            b = t;
        }
        const int old = atomicMin(&L[a], b);   // a > b: hang root a under b
        if (old == a) return;
        a = old;
    }
}

__global__ void ccl_init_kernel(int* __restrict__ L, int n) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) L[p] = p;
}

__global__ void ccl_merge_kernel(const int32_t* __restrict__ val, int* __restrict__ L, int h, int w) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= h * w) return;
    const int v = val[p];
    if (v == 0) return;
    const int x = p % w, y = p / w;
    if (x + 1 < w && val[p + 1] == v) uf_union(L, p, p + 1);
    if (y + 1 < h && val[p + w] == v) uf_union(L, p, p + w);
}

// labels[p] = 0 for background, else 1 + root index
__global__ void ccl_final_kernel(const int32_t* __restrict__ val, const int* __restrict__ L, int32_t* __restrict__ labels, int n) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    labels[p] = val[p] == 0 ? 0 : 1 + uf_find(L, p);
}

// ---------------------------------------------------------------- P2 cluster_centroids
// weak = sqrt(dx^2 + dy^2) < thres in fp32 (step/make_ins_seg_labels.py:61-64)
__global__ void weak_region_kernel(const float* __restrict__ dp, int32_t* __restrict__ weak, int n, float thres) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float a = dp[n + p], b = dp[p];
    const float s = sqrtf(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)));
    weak[p] = s < thres ? 1 : 0;
}

// value at each pixel's centroid (0 = strong-displacement bucket, else 1 + component root) and presence flags
__global__ void centroid_lookup_kernel(const int32_t* __restrict__ labels, const int32_t* __restrict__ cen, int32_t* __restrict__ value,
                                       int32_t* __restrict__ present, int h, int w) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= h * w) return;
    const int v = labels[cen[p] * w + cen[(size_t)h * w + p]];
    value[p] = v;
    present[v] = 1;
}

// exclusive prefix sum of 0/1 flags over n entries (single block); total -> *count.  rank = compress_range order.
__global__ void rank_scan_kernel(const int32_t* __restrict__ present, int32_t* __restrict__ rank, int n, int32_t* __restrict__ count) {
    __shared__ int s_part[1024];
    const int t = threadIdx.x, nt = blockDim.x;
    const int per = (n + nt - 1) / nt;
    const int b = t * per, e = min(b + per, n);
    int sum = 0;
    for (int i = b; i < e; ++i) sum += present[i] ? 1 : 0;
    s_part[t] = sum;
    __syncthreads();
    for (int o = 1; o < nt; o <<= 1) {   // Hillis-Steele inclusive scan of the partials
        const int v = t >= o ? s_part[t - o] : 0;
        __syncthreads();
        s_part[t] += v;
        __syncthreads();
    }
    int run = t ? s_part[t - 1] : 0;
    for (int i = b; i < e; ++i) {
        rank[i] = run;
        run += present[i] ? 1 : 0;
    }
    if (t == nt - 1) *count = s_part[t];
}

__global__ void instance_map_kernel(const int32_t* __restrict__ value, const int32_t* __restrict__ rank, int32_t* __restrict__ inst, int n) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) inst[p] = rank[value[p]];
}

// ---------------------------------------------------------------- P3 separte_score_by_mask
// out[(k*I + i), p] = cams[k, p] * (inst[p] == i)        (step/make_ins_seg_labels.py:77-80)
__global__ void instance_seeds_kernel(const float* __restrict__ cams, const int32_t* __restrict__ inst, float* __restrict__ out, int K,
                                      int I, int n) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)K * I * n) return;
    const int p = (int)(idx % n);
    const int c = (int)(idx / n);
    const int k = c / I, i = c % I;
    out[idx] = inst[p] == i ? cams[(size_t)k * n + p] : 0.f;
}

// ---------------------------------------------------------------- P4 detect_instance statistics
// For every segment (component of the argmax index map, id = 1 + root): area and max of the channel's score.
__global__ void segment_stats_kernel(const int32_t* __restrict__ labels, const int32_t* __restrict__ index, const float* __restrict__ scores,
                                     int32_t* __restrict__ area, int32_t* __restrict__ max_bits, int n) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int l = labels[p];
    if (l == 0) return;
    atomicAdd(&area[l], 1);
    atomicMax(&max_bits[l], __float_as_int(scores[(size_t)(index[p] - 1) * n + p]));   // scores >= 0
}

// Detection masks: out[m][p] = (labels[p] == seg_ids[m]) as bytes (numpy bool), 16 pixels per thread.
__global__ void segment_masks_kernel(const int32_t* __restrict__ labels, const int32_t* __restrict__ seg_ids, uint8_t* __restrict__ out, int M,
                                     int n) {
    const int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (p0 >= n) return;
    const int m = blockIdx.y;
    const int id = seg_ids[m];
    uint8_t v[16];
    const bool full = p0 + 16 <= n;
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (full || p0 + i < n) ? (labels[p0 + i] == id ? 1 : 0) : 0;
    uint8_t* o = out + (size_t)m * n + p0;
    if (full && (((size_t)m * n + p0) & 15) == 0) {
        *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(v);
    } else {
        for (int i = 0; i < 16 && p0 + i < n; ++i) o[i] = v[i];
    }
}

}  // namespace irn

using namespace irn;

// masks uint8 [M,H,W] (0/1) of the M segments whose labels are listed in seg_ids (device int32 [M])
extern "C" int irn_segment_masks(const int32_t* labels, const int32_t* seg_ids, int M, int H, int W, uint8_t* masks, irn_stream_t stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    launch_counter() = 0;
    if (M < 0 || H <= 0 || W <= 0) return fail(kBadArg, "irn_segment_masks: bad size");
    if (M == 0) return kOk;
    if (!labels || !seg_ids || !masks) return fail(kBadArg, "irn_segment_masks: null pointer");
    if (M > 65535) return fail(kUnsupported, "irn_segment_masks: more than 65535 segments");
    const int n = H * W;
    segment_masks_kernel<<<dim3((unsigned)((n + 16 * 256 - 1) / (16 * 256)), (unsigned)M), 256, 0, st>>>(labels, seg_ids, masks, M, n);
    IRN_LAUNCH_CHECK("segment_masks_kernel");
    return kOk;
}

extern "C" int irn_find_centroids(const float* dp, int32_t* centroids, int h, int w, int iterations, irn_stream_t stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    launch_counter() = 0;
    if (!dp || !centroids || h <= 0 || w <= 0 || iterations < 0) return fail(kBadArg, "irn_find_centroids: bad argument");
    centroids_kernel<<<(h * w + 127) / 128, 128, 0, st>>>(dp, centroids, h, w, iterations);
    IRN_LAUNCH_CHECK("centroids_kernel");
    return kOk;
}

extern "C" int irn_connected_components(const int32_t* values, int32_t* labels, int h, int w, void* scratch, irn_stream_t stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    launch_counter() = 0;
    if (!values || !labels || !scratch || h <= 0 || w <= 0) return fail(kBadArg, "irn_connected_components: bad argument");
    const int n = h * w, g = (n + 255) / 256;
    int* L = (int*)scratch;
    ccl_init_kernel<<<g, 256, 0, st>>>(L, n);
    IRN_LAUNCH_CHECK("ccl_init_kernel");
    ccl_merge_kernel<<<g, 256, 0, st>>>(values, L, h, w);
    IRN_LAUNCH_CHECK("ccl_merge_kernel");
    ccl_final_kernel<<<g, 256, 0, st>>>(values, L, labels, n);
    IRN_LAUNCH_CHECK("ccl_final_kernel");
    return kOk;
}

extern "C" size_t irn_cluster_scratch_bytes(int h, int w) { return (size_t)(5 * (size_t)h * w + 16) * sizeof(int32_t); }

extern "C" int irn_cluster_centroids(const float* dp, const int32_t* centroids, float thres, int32_t* instance_map, int32_t* n_instances_dev,
                                     int h, int w, void* scratch, irn_stream_t stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    launch_counter() = 0;
    if (!dp || !centroids || !instance_map || !n_instances_dev || !scratch || h <= 0 || w <= 0) return fail(kBadArg, "irn_cluster_centroids: bad argument");
    const int n = h * w, g = (n + 255) / 256;
    int32_t* weak = (int32_t*)scratch;
    int32_t* L = weak + n;
    int32_t* labels = L + n;
    int32_t* value = labels + n;
    int32_t* present = value + n;          // n + 1 entries
    // `rank` reuses `weak`+`L` (2n >= n + 1 entries) once the labels are final
    weak_region_kernel<<<g, 256, 0, st>>>(dp, weak, n, thres);
    IRN_LAUNCH_CHECK("weak_region_kernel");
    ccl_init_kernel<<<g, 256, 0, st>>>(L, n);
    IRN_LAUNCH_CHECK("ccl_init_kernel");
    ccl_merge_kernel<<<g, 256, 0, st>>>(weak, L, h, w);
    IRN_LAUNCH_CHECK("ccl_merge_kernel");
    ccl_final_kernel<<<g, 256, 0, st>>>(weak, L, labels, n);
    IRN_LAUNCH_CHECK("ccl_final_kernel");
    IRN_CUDA(cudaMemsetAsync(present, 0, (size_t)(n + 1) * sizeof(int32_t), st));
    centroid_lookup_kernel<<<g, 256, 0, st>>>(labels, centroids, value, present, h, w);
    IRN_LAUNCH_CHECK("centroid_lookup_kernel");
    int32_t* rank = weak;
    rank_scan_kernel<<<1, 1024, 0, st>>>(present, rank, n + 1, n_instances_dev);
    IRN_LAUNCH_CHECK("rank_scan_kernel");
    instance_map_kernel<<<g, 256, 0, st>>>(value, rank, instance_map, n);
    IRN_LAUNCH_CHECK("instance_map_kernel");
    return kOk;
}

extern "C" int irn_instance_seeds(const float* cams, const int32_t* instance_map, int K, int I, int h, int w, float* out, irn_stream_t stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    launch_counter() = 0;
    if (K < 0 || I < 0 || h <= 0 || w <= 0) return fail(kBadArg, "irn_instance_seeds: bad size");
    if (K == 0 || I == 0) return kOk;
    if (!cams || !instance_map || !out) return fail(kBadArg, "irn_instance_seeds: null pointer");
    const size_t total = (size_t)K * I * h * w;
    instance_seeds_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(cams, instance_map, out, K, I, h * w);
    IRN_LAUNCH_CHECK("instance_seeds_kernel");
    return kOk;
}

// labels int32 [H,W] from irn_connected_components(index map); area / max_bits int32 [H*W + 1] (zeroed here)
extern "C" int irn_segment_stats(const int32_t* labels, const int32_t* index, const float* scores, int H, int W, int32_t* area,
                                 int32_t* max_bits, irn_stream_t stream_) {
    cudaStream_t st = (cudaStream_t)stream_;
    launch_counter() = 0;
    if (!labels || !index || !scores || !area || !max_bits || H <= 0 || W <= 0) return fail(kBadArg, "irn_segment_stats: bad argument");
    const int n = H * W;
    IRN_CUDA(cudaMemsetAsync(area, 0, (size_t)(n + 1) * sizeof(int32_t), st));
    IRN_CUDA(cudaMemsetAsync(max_bits, 0, (size_t)(n + 1) * sizeof(int32_t), st));
    segment_stats_kernel<<<(n + 255) / 256, 256, 0, st>>>(labels, index, scores, area, max_bits, n);
    IRN_LAUNCH_CHECK("segment_stats_kernel");
    return kOk;
}
