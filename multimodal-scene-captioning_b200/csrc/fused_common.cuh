// fused_common.cuh -- types and device helpers shared by the table kernels (fused_evidence.cu) and the two streaming kernels
// (stream4.cu: config 10, the default; fused_stream.cu: config 7, the TMA-ring generation, which also serves fov_keep_mask != 0).
#pragma once
#include "msc_common.cuh"

namespace msc {

constexpr int kBoxStride = 20;      // floats per box record: 80 B stride makes the four LDS.128 conflict-free

constexpr uint32_t kCullEmpty = 0xffffffffu;  // no candidate box in this cell
constexpr uint32_t kCullAll = 0xfefefefeu;    // more than four boxes touch the cell: test every box
constexpr int kAccWords = 9;                  // per-box accumulators: count, min s, 3 axis words, 3 carry words, pad; the odd
                                              // stride spreads the same word of different boxes over all 32 banks
constexpr int kMaxWarps = 32;
constexpr int kInnerMax = 50;  // fused_stream.cu: side of the fine (one BEV cell) edge-class grid around the sensor

struct FusedLayout {  // byte offsets into dynamic smem, computed on the host
    int32_t tiles_off, window_off, cull_off, boxp_off, boxacc_off, misc_off, queue_off, total_bytes;
    int32_t win_w, win_lo;         // window covers cells [win_lo, win_lo + win_w) in x and y
    int32_t cull_dim, cull_shift;  // cull cell = BEV cell >> cull_shift
    int32_t inner_off, inner_dim, inner_lo;  // fused_stream.cu: edge classes per BEV cell for cells [inner_lo, inner_lo + inner_dim)^2 (0: none)
    int32_t max_boxes;             // capacity of the smem box tables
    int32_t pcnt_off, isum_delta;  // stream4.cu: bytes from the window region's start to the cull-cell count words / from array A to array B
    int32_t win_stride;            // stream4.cu: words between window rows (win_w + 1: odd, so rays along y do not pile up in one bank)
};

struct FusedArgs {
    msc_params P;
    msc_batch_in in;
    msc_batch_out out;
    FusedLayout L;
    // host-precomputed scalars (exact): 2r, res, RN(1/2r), 2^centroid_shift, 2^intensity_shift
    float two_r, resf, rcp_two_r, cscale, iscale;
    int32_t centroid_bias;  // 2^(centroid_shift + 6): makes the quantised coordinate non-negative
    int32_t split;          // (unused since stream3.cu was retired: always 1)
    // Replicas of the small result tables (per-box tables, projection tables, stats): the kernels that produce a value also store it
    // through these pointers -- peer-mapped (symmetric) memory of the other GPUs of the box -- so the "gather" of a sharded batch is
    // part of the producing kernels (P2P stores over NVLink / NVSwitch), not a collective after them.  bev_* members are ignored.
    int32_t n_replicas;
    msc_batch_out replica[MSC_MAX_REPLICAS];
};

// BEV cell index, lidar_agent.py:547-552.  FASTDIV replaces the IEEE division by the 3-instruction Markstein
// sequence, which tools/markstein_check.c proves equal to RN(a/b) for every float a outside the subnormal
// quotient range for the whitelisted divisors (a = fl(c + r) is 0 or >= 2^-24 r here).
template <bool FASTDIV>
__device__ __forceinline__ int bev_cell(float c, float r, float two_r, float rcp_two_r, float resf, int res_m1) {
    const float a = __fadd_rn(c, r);
    float q;
    if (FASTDIV) {
        const float q0 = __fmul_rn(a, rcp_two_r);
        const float rem = __fmaf_rn(-two_r, q0, a);
        q = __fmaf_rn(rem, rcp_two_r, q0);
    } else {
        q = __fdiv_rn(a, two_r);
    }
    const int i = __float2int_rz(__fmul_rn(q, resf));
    return min(max(i, 0), res_m1);
}

// Packed pairs of f32 (Blackwell FADD2 / FMUL2 / FFMA2: two IEEE round-to-nearest operations per lane and issue slot; each half is
// bit-identical to the scalar instruction, so the arithmetic contract of msc_common.cuh is unchanged).
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// bev_cell() of x and y in one packed pass (same operations, same order, per half)
template <bool FASTDIV>
__device__ __forceinline__ void bev_cell_xy(float x, float y, float r, float two_r, float rcp_two_r, float resf, int res_m1, int& ix, int& iy) {
    const unsigned long long a = f2_add(f2_pack(x, y), f2_pack(r, r));
    unsigned long long q;
    if (FASTDIV) {
        const unsigned long long rc = f2_pack(rcp_two_r, rcp_two_r);
        const unsigned long long q0 = f2_mul(a, rc);
        const unsigned long long rem = f2_fma(f2_pack(-two_r, -two_r), q0, a);
        q = f2_fma(rem, rc, q0);
    } else {
        float ax, ay;
        f2_unpack(a, ax, ay);
        q = f2_pack(__fdiv_rn(ax, two_r), __fdiv_rn(ay, two_r));
    }
    float tx, ty;
    f2_unpack(f2_mul(q, f2_pack(resf, resf)), tx, ty);
    ix = min(max(__float2int_rz(tx), 0), res_m1);
    iy = min(max(__float2int_rz(ty), 0), res_m1);
}

struct TableLayout {  // offsets (bytes) into the workspace
    size_t counter_off, boxprep_off, wedge_off, cullids_off;
    size_t boxscr_off, splitstats_off;  // merge scratch of samples processed in parts: per-box (count | min << 32, 3 biased sums) u64 x 4, per-sample stats + ticket
    size_t tileoff_off;                 // stream4.cu: [n_samples + 1] exclusive prefix of warp tiles per sample
    size_t cls_off;                     // stream4.cu: per-CTA class words, [max CTAs][cull_dim^2 + kInnerMax^2] u32
    size_t total;
};

// fused_stream.cu (config 7): bytes of its per-CTA state block, and its launcher
int stream_misc_bytes();
void stream_shape_info(int* threads, int* tile_pts, int* ring_bytes, int* queue_bytes);
int launch_stream_kernel(const FusedArgs& args, const TableLayout& T, unsigned char* ws, int grid, bool fov, bool fast, cudaStream_t stream);
// stream4.cu (config 10)
int stream4_threads(int ppt);
FusedLayout stream4_layout(int smem_bytes, int ppt, int res, int cull_dim, int cull_shift, int box_cap, int opt_window, int inner_dim);
bool stream4_is_standard(const FusedArgs& a, int smem_bytes, int ppt, int opt_window);
int stream4_std_boxes();
int launch_stream4_partition(const FusedArgs& args, const TableLayout& T, unsigned char* ws, int grid, cudaStream_t stream, int* launches);
int launch_stream4_kernel(const FusedArgs& args, const TableLayout& T, unsigned char* ws, int grid, int ppt, bool fov, bool fast, bool standard,
                          cudaStream_t stream);

// Per-edge classes of one cell against one wedge: bit0 = the wedge may contain points of the cell,
// bit1 = the right edge is undecided inside the cell, bit2 = the left edge is.  Same extremes over the cell and the same guard band
// as classify_cell(); an edge every point of the cell passes needs no exact test, an edge every point fails empties the wedge.
__device__ __forceinline__ uint32_t classify_cell_edges(const float* __restrict__ wq, float x0, float x1, float y0, float y1) {
    // both cross products are affine in (x, y): over the rectangle they range over (value at the centre) -/+ (extent), which costs a
    // third of four corner evaluations.  The float evaluation error of the exact test is < 1e-4 for |p| < 128 m and << guard for |p|
    // up to the 1200 m "absorbing" edge cells, far inside the 2e-3 guard band.
    const float guard = 2e-3f;
    const float hx = 0.5f * (x1 - x0), hy = 0.5f * (y1 - y0);
    const float qx = 0.5f * (x0 + x1) - wq[0], qy = 0.5f * (y0 + y1) - wq[1];
    const float cr = wq[4] * qy - wq[5] * qx, cr_e = fabsf(wq[4]) * hy + fabsf(wq[5]) * hx;
    const float cl = qx * wq[3] - qy * wq[2], cl_e = fabsf(wq[3]) * hx + fabsf(wq[2]) * hy;
    if (cr + cr_e < -guard || cl + cl_e < -guard) return 0u;  // outside
    return 1u | ((cr - cr_e > guard) ? 0u : 2u) | ((cl - cl_e > guard) ? 0u : 4u);
}

// Class word of entry i of a sample's class tables (i < cull_dim^2: cull cell i; beyond: fine cell i - cull_dim^2, one BEV cell of the
// inner_dim x inner_dim square around the sensor), from the sample's camera wedges wq[MSC_MAX_CAMS][6] (any address space):
// bit c = camera c's wedge may contain points of the cell, bit 8 + c = its right edge is undecided inside the cell, bit 16 + c = its left
// edge is.  Every point whose bev_cell() index falls in the cell lies in the padded rectangle; first / last cells absorb what is clipped
// into them.  Both streaming kernels evaluate this per sample in their prologue (5,000 cells x 6 cameras: < 1 % of a sample's work).
__device__ __forceinline__ uint32_t edge_class_word(const FusedArgs& A, const float* __restrict__ wq, int i) {
    const msc_params& P = A.P;
    const int ncc = A.L.cull_dim * A.L.cull_dim;
    const float big = 4.0f * P.bev_range + 1000.0f, pad = 2e-3f;
    float x0, x1, y0, y1;
    if (i >= ncc) {
        const int j = i - ncc, jy = j / A.L.inner_dim, jx = j - jy * A.L.inner_dim;
        const float cell_b = A.two_r / A.resf;
        const int ix = A.L.inner_lo + jx, iy = A.L.inner_lo + jy, last_b = P.bev_res - 1;
        x0 = (ix == 0) ? -big : (-P.bev_range + (float)ix * cell_b - pad); x1 = (ix == last_b) ? big : (-P.bev_range + (float)(ix + 1) * cell_b + pad);
        y0 = (iy == 0) ? -big : (-P.bev_range + (float)iy * cell_b - pad); y1 = (iy == last_b) ? big : (-P.bev_range + (float)(iy + 1) * cell_b + pad);
    } else {
        const int gy = i / A.L.cull_dim, gx = i - gy * A.L.cull_dim;
        const float cell_m = (A.two_r / A.resf) * (float)(1 << A.L.cull_shift);
        const int last = A.L.cull_dim - 1;
        x0 = (gx == 0) ? -big : (-P.bev_range + (float)gx * cell_m - pad); x1 = (gx == last) ? big : (-P.bev_range + (float)(gx + 1) * cell_m + pad);
        y0 = (gy == 0) ? -big : (-P.bev_range + (float)gy * cell_m - pad); y1 = (gy == last) ? big : (-P.bev_range + (float)(gy + 1) * cell_m + pad);
    }
    uint32_t eb = 0;
    for (int c = 0; c < P.n_cams; ++c) {
        const uint32_t k = classify_cell_edges(wq + c * 6, x0, x1, y0, y1);
        eb |= ((k & 1u) << c) | (((k >> 1) & 1u) << (8 + c)) | (((k >> 2) & 1u) << (16 + c));  // in-bit, right / left edge undecided
    }
    return eb;
}

// exact wedge test (the definition): q = p - apex, cross(e_right, q) >= 0 and cross(q, e_left) >= 0
__device__ __forceinline__ bool in_wedge(const float* __restrict__ wq, float x, float y) {
    const float qx = __fsub_rn(x, wq[0]), qy = __fsub_rn(y, wq[1]);
    const float cr = __fmaf_rn(wq[4], qy, -__fmul_rn(wq[5], qx));
    const float cl = __fmaf_rn(-wq[2], qy, __fmul_rn(wq[3], qx));  // same shape as cr: product on the x term, fma on the y term
    return (cr >= 0.0f) && (cl >= 0.0f);
}

// Conservative oriented rasterisation of a prepared box's xy footprint (a zonotope spanned by the projected edge
// vectors) into the cull grid.  Every member point lies in the corner hull up to float rounding (<< the 2 mm
// margin) and bev_cell() is monotonic, so a member can never fall in an unmarked cell.
// cull_words: first id word of cell 0; STRIDE words between cells (2: the smem (ids, classes) pairs, 1: the workspace id table).
// The cells of the box's bounding rectangle are dealt round-robin to `nworkers` callers (worker = 0 .. nworkers - 1).
template <int STRIDE>
static __device__ __noinline__ void rasterise_box(const FusedArgs& A, const float* __restrict__ o, int b, uint32_t* __restrict__ cull_words,
                                                  int worker = 0, int nworkers = 1) {
    const msc_params& P = A.P;
    const FusedLayout& L = A.L;
    const float margin = 2e-3f;
    const float cx = o[16], cy = o[17];
    const float ex[3] = {o[3], o[6], o[9]}, ey[3] = {o[4], o[7], o[10]};
    const float rx = 0.5f * (fabsf(ex[0]) + fabsf(ex[1]) + fabsf(ex[2])) + margin;
    const float ry = 0.5f * (fabsf(ey[0]) + fabsf(ey[1]) + fabsf(ey[2])) + margin;
    const int res_m1 = P.bev_res - 1;
    const int cx0 = bev_cell<false>(cx - rx, P.bev_range, A.two_r, A.rcp_two_r, A.resf, res_m1) >> L.cull_shift;
    const int cx1 = bev_cell<false>(cx + rx, P.bev_range, A.two_r, A.rcp_two_r, A.resf, res_m1) >> L.cull_shift;
    const int cy0 = bev_cell<false>(cy - ry, P.bev_range, A.two_r, A.rcp_two_r, A.resf, res_m1) >> L.cull_shift;
    const int cy1 = bev_cell<false>(cy + ry, P.bev_range, A.two_r, A.rcp_two_r, A.resf, res_m1) >> L.cull_shift;
    const float cell_m = (A.two_r / A.resf) * (float)(1 << L.cull_shift);
    const int last = L.cull_dim - 1;
    const int nx_cells = cx1 - cx0 + 1, n_cells = nx_cells * (cy1 - cy0 + 1);
    for (int ci = worker; ci < n_cells; ci += nworkers) {
        {
            const int gy = cy0 + ci / nx_cells, gx = cx0 + ci % nx_cells;
            // edge cells absorb everything clipped into them: never reject those
            bool reject = false;
            if (gx > 0 && gx < last && gy > 0 && gy < last) {
                const float mx = -P.bev_range + ((float)gx + 0.5f) * cell_m, my = -P.bev_range + ((float)gy + 0.5f) * cell_m;
                const float dx = cx - mx, dy = cy - my;
                const float hc = 0.5f * cell_m + margin;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float nx = -ey[k], ny = ex[k];  // normal of projected edge k
                    const float nn = fabsf(nx) + fabsf(ny);
                    if (nn > 1e-6f) {
                        const float dist = fabsf(dx * nx + dy * ny);
                        float rb = 0.0f;
#pragma unroll
                        for (int j = 0; j < 3; ++j) rb += 0.5f * fabsf(ex[j] * nx + ey[j] * ny);
                        if (dist > (rb + hc * nn) * 1.0001f + margin * nn) reject = true;
                    }
                }
            }
            if (reject) continue;
            uint32_t* slot = cull_words + (size_t)(gy * L.cull_dim + gx) * STRIDE;
            for (;;) {
                const uint32_t old = *reinterpret_cast<volatile uint32_t*>(slot);
                if (old == kCullAll) break;
                uint32_t nw = kCullAll;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (nw == kCullAll && ((old >> (8 * k)) & 0xffu) == 0xffu) nw = (old & ~(0xffu << (8 * k))) | ((uint32_t)b << (8 * k));
                if (atomicCAS(slot, old, nw) == old) break;
            }
        }
    }
}

// A.2 exact membership test of one point against one prepared box (closed intervals, float32 with fmaf chains).
__device__ __forceinline__ bool box_contains(const float* __restrict__ boxp, int b, float xr, float yr, float zr) {
    const float4* bp = reinterpret_cast<const float4*>(boxp + b * kBoxStride);
    const float4 b0 = bp[0], b1 = bp[1], b2 = bp[2], b3 = bp[3];
    const float v0 = __fsub_rn(xr, b0.x), v1 = __fsub_rn(yr, b0.y), v2 = __fsub_rn(zr, b0.z);
    const float iv = __fmaf_rn(b1.y, v2, __fmaf_rn(b1.x, v1, __fmul_rn(b0.w, v0)));
    const float jv = __fmaf_rn(b2.x, v2, __fmaf_rn(b1.w, v1, __fmul_rn(b1.z, v0)));
    const float kv = __fmaf_rn(b2.w, v2, __fmaf_rn(b2.z, v1, __fmul_rn(b2.y, v0)));
    return iv >= 0.0f && iv <= b3.x && jv >= 0.0f && jv <= b3.y && kv >= 0.0f && kv <= b3.z;
}
}  // namespace msc
