// fused_evidence.cu -- the hot path: one pass over the raw sweeps of a batch of samples.
//
// Per raw point (20 B as stored in .pcd.bin): remove_close -> f64 rigid transform of its sweep ->
// range/height filter -> camera-wedge (FOV) membership -> ground/object split -> BEV cell update
// (count, Q8 intensity sum, max height) -> oriented-box membership through a per-sample cull grid
// (count, nearest, fixed-point centroid sums).  Per sample: box preparation (global -> ego -> sensor),
// cull-grid rasterisation, camera wedges and box->camera projection in the prologue; result tables in
// the epilogue.  Semantics: SURVEY.md App. A + lidar_agent.py:103-132, :547-560 (cited per step below).
//
// B200 mapping:
//   * persistent grid, one CTA per SM, one sample per CTA at a time (dynamic work counter);
//   * raw sweep rows stream HBM -> smem through a ring of 20 KB cp.async.bulk (TMA) tiles with mbarrier
//     completion and an L2 evict-first policy; threads read x,y,z,i at a 5-word stride, which is
//     bank-conflict free (5 is odd);
//   * the BEV accumulators of a centred window of the grid live in smem as (count u32, isum u32) pairs
//     updated with native integer ATOMS; cells outside the window take ONE 64-bit RED on the interleaved
//     global cell; the window is flushed once per sample with coalesced 16-byte stores;
//   * one 8-byte smem entry per 2 m cull cell carries up to four candidate box ids (oriented, conservative
//     rasterisation) and the per-camera wedge classification (inside / straddling), so a point costs one
//     LDS.64 to learn which boxes and which exact wedge tests it needs;
//   * per-box accumulators are smem-resident integers: count, min s, and centroid sums split into three
//     9-bit limbs per axis so every update is a fire-and-forget ATOMS (order-independent, bit-reproducible);
//   * no tensor cores: nothing here is a contraction.
#include "fused_common.cuh"

namespace msc {

// Launch shape: NT threads = NT/32 warps; every warp owns a private ring of STAGES tiles of 32*PPT points.
template <int NT, int PPT, int STAGES, bool POSE_SMEM = false, int QUEUE = 0>
struct Cfg {
    // QUEUE = 1: per-warp queue of points that have candidate boxes, drained 32 at a time so every lane tests a real candidate
    // (deferring the exact wedge tests of straddling cells the same way was measured slower and is not implemented)
    static constexpr int kQueue = QUEUE;
    static constexpr int kQueueBytes = QUEUE ? (NT / 32) * 64 * 16 : 0;
    static constexpr bool kPoseInSmem = POSE_SMEM;  // pose rows read from smem per tile (64-register budgets)
    static constexpr int kThreads = NT;
    static constexpr int kWarps = NT / 32;
    static constexpr int kPtsPerThread = PPT;
    static constexpr int kTilePts = 32 * PPT;          // points per warp tile
    static constexpr int kTileBytes = kTilePts * 20;   // 1280 B for PPT = 2
    static constexpr int kStages = STAGES;
    static constexpr int kRingBytes = kWarps * STAGES * kTileBytes;
    static_assert(kTileBytes % 16 == 0, "bulk copies move multiples of 16 bytes");
    static_assert(STAGES >= 2 && STAGES <= 8, "ring depth");
};

struct Misc {  // small per-CTA state at misc_off
    uint64_t full_bar[kMaxWarps * 8];  // [warp][stage]: TMA bytes landed in that warp's ring slot
    float wedge[MSC_MAX_CAMS][6];
    uint32_t stats[MSC_STATS_STRIDE];
    uint32_t sweep_start[kMaxSweepsSmem], sweep_count[kMaxSweepsSmem];
    int32_t sample;
    int32_t pad_[3];
    double pose[kMaxSweepsSmem * 12];  // this sample's 3x4 sweep transforms (only used by POSE_SMEM shapes)
};

// volatile so the compiler can neither rematerialise nor re-issue the load: the pose stays in registers
__device__ __forceinline__ void ld_pose(const double* __restrict__ p, double M[12]) {
#pragma unroll
    for (int i = 0; i < 12; i += 2)
        asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(M[i]), "=d"(M[i + 1]) : "l"(p + i));
}

// ------------------------------------------------------------------------------------------------ table kernel
// Everything that does not touch points runs once per batch in a small, fully parallel kernel and lands in the
// workspace: prepared boxes (devkit points_in_box vectors, App. A.2), box -> camera projection (A.3), camera
// wedges, and the per-cull-cell wedge classes.  The streaming kernel then only copies its sample's rows to smem.


// classify one cull cell against one wedge: bit0 = every point of the cell is inside, bit1 = undecided
__device__ __forceinline__ uint32_t classify_cell(const float* __restrict__ wq, float x0, float x1, float y0, float y1) {
    // both cross products are affine in (x, y): extremes over the rectangle are at its corners.  The float
    // evaluation error of the exact test is < 1e-4 for |p| < 128 m, far inside the 2e-3 guard band.
    const float guard = 2e-3f;
    float cr_min = INFINITY, cr_max = -INFINITY, cl_min = INFINITY, cl_max = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float qx = ((k & 1) ? x1 : x0) - wq[0], qy = ((k & 2) ? y1 : y0) - wq[1];
        const float cr = wq[4] * qy - wq[5] * qx, cl = qx * wq[3] - qy * wq[2];
        cr_min = fminf(cr_min, cr); cr_max = fmaxf(cr_max, cr);
        cl_min = fminf(cl_min, cl); cl_max = fmaxf(cl_max, cl);
    }
    if (cr_min > guard && cl_min > guard) return 1u;    // inside
    if (cr_max < -guard || cl_max < -guard) return 0u;  // outside
    return 2u;                                          // straddling: run the exact test per point
}

// Per-edge classes of one cull cell against one wedge (fused_stream.cu): bit0 = the wedge may contain points of the cell,
// bit1 = the right edge is undecided inside the cell, bit2 = the left edge is.  Same extremes over the cell and the same guard band
// as classify_cell(); an edge every point of the cell passes needs no exact test, an edge every point fails empties the wedge.
__device__ __forceinline__ uint32_t classify_cell_edges(const float* __restrict__ wq, float x0, float x1, float y0, float y1) {
    // both cross products are affine in (x, y): over the rectangle they range over (value at the centre) -/+ (extent), which costs a
    // third of four corner evaluations.  Float error << guard for |p| up to the 1200 m "absorbing" edge cells.
    const float guard = 2e-3f;
    const float hx = 0.5f * (x1 - x0), hy = 0.5f * (y1 - y0);
    const float qx = 0.5f * (x0 + x1) - wq[0], qy = 0.5f * (y0 + y1) - wq[1];
    const float cr = wq[4] * qy - wq[5] * qx, cr_e = fabsf(wq[4]) * hy + fabsf(wq[5]) * hx;
    const float cl = qx * wq[3] - qy * wq[2], cl_e = fabsf(wq[3]) * hx + fabsf(wq[2]) * hy;
    if (cr + cr_e < -guard || cl + cl_e < -guard) return 0u;  // outside
    return 1u | ((cr - cr_e > guard) ? 0u : 2u) | ((cl - cl_e > guard) ? 0u : 4u);
}

// camera wedge: apex = camera centre in the sensor xy-plane, edges = image columns 0 and W (f64, no FMA)
__device__ void compute_wedge(int image_w, const double* __restrict__ lcal, const double* __restrict__ ccal, const double* __restrict__ K,
                              float* __restrict__ wq) {
    double Rl[9], Rc[9];
    quat_to_rot(lcal + 3, Rl);
    quat_to_rot(ccal + 3, Rc);
    const double d0 = ccal[0] - lcal[0], d1 = ccal[1] - lcal[1], d2 = ccal[2] - lcal[2];
    const double ox = (Rl[0] * d0 + Rl[3] * d1) + Rl[6] * d2;
    const double oy = (Rl[1] * d0 + Rl[4] * d1) + Rl[7] * d2;
    const double fx = K[0], cxp = K[2];
    const double dl[3] = {(0.0 - cxp) / fx, 0.0, 1.0};
    const double dr[3] = {((double)image_w - cxp) / fx, 0.0, 1.0};
    double le[3], re[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        le[r] = (Rc[r * 3 + 0] * dl[0] + Rc[r * 3 + 1] * dl[1]) + Rc[r * 3 + 2] * dl[2];
        re[r] = (Rc[r * 3 + 0] * dr[0] + Rc[r * 3 + 1] * dr[1]) + Rc[r * 3 + 2] * dr[2];
    }
    wq[0] = (float)ox; wq[1] = (float)oy;
    wq[2] = (float)((Rl[0] * le[0] + Rl[3] * le[1]) + Rl[6] * le[2]);
    wq[3] = (float)((Rl[1] * le[0] + Rl[4] * le[1]) + Rl[7] * le[2]);
    wq[4] = (float)((Rl[0] * re[0] + Rl[3] * re[1]) + Rl[6] * re[2]);
    wq[5] = (float)((Rl[1] * re[0] + Rl[4] * re[1]) + Rl[7] * re[2]);
}

// grid: ceil(max(n_boxes_total, n_samples) * max(n_cams,1) / 256) blocks of 256 threads
__global__ void __launch_bounds__(256) fused_tables_kernel(const __grid_constant__ FusedArgs A, const TableLayout T, int n_boxes_total,
                                                          unsigned char* __restrict__ ws) {
    const msc_params& P = A.P;
    const int n_cams = P.n_cams;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    float* const boxprep = reinterpret_cast<float*>(ws + T.boxprep_off);
    float* const wedges = reinterpret_cast<float*>(ws + T.wedge_off);
    uint16_t* const fovcls = reinterpret_cast<uint16_t*>(ws + T.fovcls_off);
    // sample of a global box index: binary search in sample_box_off
    auto sample_of_box = [&](int gb) {
        int lo = 0, hi = A.in.n_samples;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (A.in.sample_box_off[mid] <= gb) lo = mid; else hi = mid; }
        return lo;
    };
    // (1) prepared boxes: one thread per box
    if (gid < n_boxes_total) {
        const int sample = sample_of_box(gid);
        const double* box = A.in.boxes + (size_t)gid * 10;
        double c[3] = {box[0], box[1], box[2]};
        double R[9];
        quat_to_rot(box + 6, R);
        frame_change(A.in.ego_pose + (size_t)sample * 7, c, R);
        frame_change(A.in.lidar_calib + (size_t)sample * 7, c, R);
        const double w = box[3], l = box[4], h = box[5];
        const double hl = l / 2.0, hw = w / 2.0, hh = h / 2.0;
        float* o = boxprep + (size_t)gid * kBoxStride;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            o[r] = (float)(((R[r * 3 + 0] * hl + R[r * 3 + 1] * hw) + R[r * 3 + 2] * hh) + c[r]);
            o[3 + r] = (float)(-(l * R[r * 3 + 0]));
            o[6 + r] = (float)(-(w * R[r * 3 + 1]));
            o[9 + r] = (float)(-(h * R[r * 3 + 2]));
        }
        o[12] = __fmaf_rn(o[5], o[5], __fmaf_rn(o[4], o[4], __fmul_rn(o[3], o[3])));
        o[13] = __fmaf_rn(o[8], o[8], __fmaf_rn(o[7], o[7], __fmul_rn(o[6], o[6])));
        o[14] = __fmaf_rn(o[11], o[11], __fmaf_rn(o[10], o[10], __fmul_rn(o[9], o[9])));
        o[15] = 0.0f;
        o[16] = (float)c[0]; o[17] = (float)c[1]; o[18] = (float)c[2]; o[19] = 0.0f;  // centre, for the cull rasterisation
    }
    // (2) box -> camera projection (App. A.3): one thread per (box, camera)
    if (n_cams > 0 && gid < n_boxes_total * n_cams) {
        const int gb = gid / n_cams, c = gid - gb * n_cams;
        const int sample = sample_of_box(gb);
        project_box(A.in.boxes + (size_t)gb * 10, A.in.cam_ego_pose + ((size_t)sample * n_cams + c) * 7,
                    A.in.cam_calib + ((size_t)sample * n_cams + c) * 7, A.in.cam_K + ((size_t)sample * n_cams + c) * 9, (double)P.image_w,
                    (double)P.image_h, A.out.proj_visible + gid, A.out.proj_extent + (size_t)gid * 4);
    }
    // (3) camera wedges: one thread per (sample, camera)
    if (n_cams > 0 && gid < A.in.n_samples * n_cams) {
        const int sample = gid / n_cams, c = gid - sample * n_cams;
        compute_wedge(P.image_w, A.in.lidar_calib + (size_t)sample * 7, A.in.cam_calib + ((size_t)sample * n_cams + c) * 7,
                      A.in.cam_K + ((size_t)sample * n_cams + c) * 9, wedges + ((size_t)sample * MSC_MAX_CAMS + c) * 6);
    }
}

// wedge classes per (sample, cull cell), from the wedges the table kernel wrote (launched after it on the same stream)
// per_edge = false: wedge classes (u16) for the first-generation kernel; true: edge classes (u32) for fused_stream.cu
__global__ void __launch_bounds__(256) fused_fovcls_kernel(const __grid_constant__ FusedArgs A, const TableLayout T, unsigned char* __restrict__ ws,
                                                          bool per_edge) {
    const msc_params& P = A.P;
    const int n_cams = P.n_cams;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const float* const wedges = reinterpret_cast<const float*>(ws + T.wedge_off);
    uint16_t* const fovcls = reinterpret_cast<uint16_t*>(ws + T.fovcls_off);
    const int ncc = A.L.cull_dim * A.L.cull_dim;
    const int n_inner = per_edge ? A.L.inner_dim * A.L.inner_dim : 0;  // fine cells around the sensor (fused_stream.cu only)
    const int per_sample = ncc + n_inner;
    if (gid >= A.in.n_samples * per_sample) return;
    const int sample = gid / per_sample, i = gid - sample * per_sample;
    const float big = 4.0f * P.bev_range + 1000.0f, pad = 2e-3f;
    if (i >= ncc) {  // one BEV cell: every point whose bev_cell() index is (ix, iy) lies in the padded square
        const int j = i - ncc, jy = j / A.L.inner_dim, jx = j - jy * A.L.inner_dim;
        const float cell_b = A.two_r / A.resf;
        const int ix = A.L.inner_lo + jx, iy = A.L.inner_lo + jy, last_b = P.bev_res - 1;  // first / last BEV cells absorb what is clipped into them
        const float fx0 = (ix == 0) ? -big : (-P.bev_range + (float)ix * cell_b - pad), fx1 = (ix == last_b) ? big : (-P.bev_range + (float)(ix + 1) * cell_b + pad);
        const float fy0 = (iy == 0) ? -big : (-P.bev_range + (float)iy * cell_b - pad), fy1 = (iy == last_b) ? big : (-P.bev_range + (float)(iy + 1) * cell_b + pad);
        uint32_t eb = 0;
        for (int c = 0; c < n_cams; ++c) {
            const uint32_t k = classify_cell_edges(wedges + ((size_t)sample * MSC_MAX_CAMS + c) * 6, fx0, fx1, fy0, fy1);
            eb |= ((k & 1u) << c) | (((k >> 1) & 1u) << (8 + c)) | (((k >> 2) & 1u) << (16 + c));
        }
        reinterpret_cast<uint32_t*>(ws + T.innercls_off)[(size_t)sample * n_inner + j] = eb;
        return;
    }
    const int gy = i / A.L.cull_dim, gx = i - gy * A.L.cull_dim;
    const float cell_m = (A.two_r / A.resf) * (float)(1 << A.L.cull_shift);
    const int last = A.L.cull_dim - 1;
    // edge cells absorb everything clipped into them
    const float x0 = (gx == 0) ? -big : (-P.bev_range + (float)gx * cell_m - pad);
    const float x1 = (gx == last) ? big : (-P.bev_range + (float)(gx + 1) * cell_m + pad);
    const float y0 = (gy == 0) ? -big : (-P.bev_range + (float)gy * cell_m - pad);
    const float y1 = (gy == last) ? big : (-P.bev_range + (float)(gy + 1) * cell_m + pad);
    uint32_t bits = 0, ebits = 0;
    if (per_edge) {
        for (int c = 0; c < n_cams; ++c) {
            const uint32_t e = classify_cell_edges(wedges + ((size_t)sample * MSC_MAX_CAMS + c) * 6, x0, x1, y0, y1);
            ebits |= ((e & 1u) << c) | (((e >> 1) & 1u) << (8 + c)) | (((e >> 2) & 1u) << (16 + c));  // in-bit, right / left edge undecided
        }
        reinterpret_cast<uint32_t*>(ws + T.edgecls_off)[(size_t)sample * ncc + i] = ebits;
    } else {
        for (int c = 0; c < n_cams; ++c) {
            const uint32_t k = classify_cell(wedges + ((size_t)sample * MSC_MAX_CAMS + c) * 6, x0, x1, y0, y1);
            bits |= ((k & 1u) << c) | (((k >> 1) & 1u) << (8 + c));
        }
        fovcls[(size_t)sample * ncc + i] = (uint16_t)bits;
    }
}

// Candidate-box ids per (sample, cull cell) for fused_stream.cu: the same conservative rasterisation the first-generation kernel
// runs inside its per-sample prologue, one warp per box (lanes share the cells of its bounding rectangle), into a workspace table
// the host pre-fills with kCullEmpty.
__global__ void __launch_bounds__(128) fused_cullids_kernel(const __grid_constant__ FusedArgs A, const TableLayout T, unsigned char* __restrict__ ws) {
    const int sample = blockIdx.x, lane = threadIdx.x & 31;
    const int b = (blockIdx.y * blockDim.x + threadIdx.x) >> 5;  // box of this warp inside its sample
    const int bx0 = A.in.sample_box_off[sample];
    int n_boxes = A.in.sample_box_off[sample + 1] - bx0;
    if (n_boxes > A.L.max_boxes) n_boxes = A.L.max_boxes;  // caller under-declared max_boxes_per_sample: the streaming kernel drops these boxes too
    if (b >= n_boxes) return;
    const float* o = reinterpret_cast<const float*>(ws + T.boxprep_off) + (size_t)(bx0 + b) * kBoxStride;
    uint32_t* ids = reinterpret_cast<uint32_t*>(ws + T.cullids_off) + (size_t)sample * (size_t)(A.L.cull_dim * A.L.cull_dim);
    rasterise_box<1>(A, o, b, ids, lane, 32);
}

// ------------------------------------------------------------------------------------------------ streaming kernel

template <class C, bool FOV, bool FASTDIV>
__global__ void __launch_bounds__(C::kThreads, 1) fused_evidence_kernel(const __grid_constant__ FusedArgs A, const TableLayout T,
                                                                       unsigned char* __restrict__ ws) {
    constexpr int NT = C::kThreads, S = C::kStages, TP = C::kTilePts, PPT = C::kPtsPerThread, W = C::kWarps;
    constexpr bool MSMEM = C::kPoseInSmem;
    extern __shared__ __align__(128) unsigned char smem[];
    const msc_params& P = A.P;
    const FusedLayout& L = A.L;
    uint2* const window = reinterpret_cast<uint2*>(smem + L.window_off);
    uint2* const cull = reinterpret_cast<uint2*>(smem + L.cull_off);            // .x box ids, .y wedge classes
    float* const boxp = reinterpret_cast<float*>(smem + L.boxp_off);            // [max_boxes][kBoxStride]
    uint32_t* const boxacc = reinterpret_cast<uint32_t*>(smem + L.boxacc_off);  // [max_boxes][kAccWords]
    Misc* const misc = reinterpret_cast<Misc*>(smem + L.misc_off);
    uint32_t* const work_counter = reinterpret_cast<uint32_t*>(ws + T.counter_off);
    const float* const g_boxprep = reinterpret_cast<const float*>(ws + T.boxprep_off);
    const float* const g_wedges = reinterpret_cast<const float*>(ws + T.wedge_off);
    const uint16_t* const g_fovcls = reinterpret_cast<const uint16_t*>(ws + T.fovcls_off);

    const int tid = threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* const ring = reinterpret_cast<float*>(smem + L.tiles_off) + (size_t)warp * S * (C::kTileBytes / 4);  // this warp's slots
    uint64_t* const full = misc->full_bar + warp * 8;
    float4* const queue = reinterpret_cast<float4*>(smem + L.queue_off) + warp * 64;  // this warp's ring of pending candidate points
    const int res = P.bev_res, res_m1 = P.bev_res - 1;
    const size_t ncell = (size_t)res * (size_t)res;
    const int n_cams = P.n_cams;
    const uint64_t policy = l2_policy_evict_first();

    if (lane == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    uint32_t wk = 0;  // tiles this warp has consumed since launch (ring position and mbarrier parity)
    __syncthreads();

    for (;;) {
        // ------------------------------------------------------------ fetch a sample
        if (tid == 0) misc->sample = (int32_t)atomicAdd(work_counter, 1u);
        __syncthreads();
        const int sample = misc->sample;
        if (sample >= A.in.n_samples) break;

        const int sw0 = A.in.sample_sweep_off[sample], sw1 = A.in.sample_sweep_off[sample + 1];
        const int n_sw = sw1 - sw0;
        if (tid < kMaxSweepsSmem && tid < n_sw) {
            misc->sweep_start[tid] = A.in.sweep_start[sw0 + tid];
            misc->sweep_count[tid] = A.in.sweep_count[sw0 + tid];
        }
        if (MSMEM) {
            for (int i = tid; i < min(n_sw, kMaxSweepsSmem) * 12; i += NT) misc->pose[i] = A.in.sweep_pose[(size_t)sw0 * 12 + i];
        }
        if (tid < MSC_STATS_STRIDE) misc->stats[tid] = 0u;
        __syncthreads();
        auto sweep_cnt = [&](int si) -> uint32_t { return si < kMaxSweepsSmem ? misc->sweep_count[si] : A.in.sweep_count[sw0 + si]; };
        auto sweep_beg = [&](int si) -> uint32_t { return si < kMaxSweepsSmem ? misc->sweep_start[si] : A.in.sweep_start[sw0 + si]; };

        // Warp `warp` owns tiles warp, warp + W, warp + 2W, ... of every sweep.  Two cursors walk that sequence:
        // p_* for the refills (S - 1 tiles ahead) and c_* for consumption; both cache their sweep's size and base.
        int p_si = -1, c_si = -1;
        uint32_t p_first = 0, p_cnt = 0, p_base = 0, c_first = 0, c_cnt = 0, p_issued = 0;
        const uint32_t wk0 = wk;
        auto issue_next = [&]() {  // whole warp (uniform control flow); lane 0 talks to the TMA unit
            while (p_first >= p_cnt) {  // next sweep that still has a tile for this warp
                if (++p_si >= n_sw) { p_si = n_sw; p_cnt = 0; p_first = 0; return; }
                p_cnt = sweep_cnt(p_si); p_base = sweep_beg(p_si); p_first = (uint32_t)warp * TP;
            }
            const uint32_t npts = min((uint32_t)TP, p_cnt - p_first);
            const int stage = (int)((wk0 + p_issued) % S);
            if (lane == 0) {
                const uint32_t bytes = (npts * 20u + 15u) & ~15u;
                mbar_arrive_expect_tx(&full[stage], bytes);
                bulk_load(ring + (size_t)stage * (C::kTileBytes / 4), A.in.points + ((size_t)p_base + p_first) * 5, bytes, &full[stage], policy);
            }
            p_first += W * TP;
            ++p_issued;
        };
#pragma unroll 1
        for (int s = 0; s < S - 1; ++s) issue_next();  // these loads overlap the prologue below

        // ------------------------------------------------------------ prologue: zero accumulators, copy tables to smem
        const int bx0 = A.in.sample_box_off[sample];
        int n_boxes = A.in.sample_box_off[sample + 1] - bx0;
        const bool box_overflow = n_boxes > L.max_boxes;  // caller under-declared max_boxes_per_sample
        if (box_overflow) n_boxes = L.max_boxes;
        uint32_t* const g_ci = A.out.bev_ci + (size_t)sample * ncell * 2;
        float* const g_h = A.out.bev_height + (size_t)sample * ncell;
        {
            uint4* w4 = reinterpret_cast<uint4*>(window);
            const int n_w4 = (L.win_w * L.win_w * 8) / 16;
            for (int i = tid; i < n_w4; i += NT) w4[i] = make_uint4(0, 0, 0, 0);
            const int n_cull = L.cull_dim * L.cull_dim;
            const uint16_t* fc = g_fovcls + (size_t)sample * n_cull;
            for (int i = tid; i < n_cull; i += NT) cull[i] = make_uint2(kCullEmpty, (FOV && n_cams > 0) ? (uint32_t)fc[i] : 0u);
            for (int i = tid; i < n_boxes * kAccWords; i += NT) boxacc[i] = ((i % kAccWords) == 1) ? 0x7f800000u : 0u;
            const float4* bsrc = reinterpret_cast<const float4*>(g_boxprep + (size_t)bx0 * kBoxStride);
            for (int i = tid; i < n_boxes * (kBoxStride / 4); i += NT) reinterpret_cast<float4*>(boxp)[i] = bsrc[i];
            if (FOV && tid < n_cams * 6) misc->wedge[tid / 6][tid % 6] = g_wedges[((size_t)sample * MSC_MAX_CAMS + tid / 6) * 6 + tid % 6];
            // zero-fill this sample's global layers (window cells are overwritten by the flush; filling them too
            // keeps the stores fully coalesced)
            uint4* c4 = reinterpret_cast<uint4*>(g_ci);
            for (size_t i = tid; i < ncell / 2; i += NT) c4[i] = make_uint4(0, 0, 0, 0);
            uint4* h4 = reinterpret_cast<uint4*>(g_h);
            for (size_t i = tid; i < ncell / 4; i += NT) h4[i] = make_uint4(0, 0, 0, 0);
        }
        __threadfence();
        __syncthreads();
        for (int b = tid; b < n_boxes; b += NT) rasterise_box<2>(A, boxp + b * kBoxStride, b, reinterpret_cast<uint32_t*>(cull));
        __syncthreads();

        // ------------------------------------------------------------ main loop: this warp's tiles, no cross-warp sync
        uint32_t c_close = 0, c_kept = 0, c_ground = 0;  // per-thread counters (flushed once per sample)
        uint32_t cam_lo = 0, cam_hi = 0;                 // eight 8-bit per-camera counters, spilled every <= 255 points
        uint32_t cam_pts = 0;
        uint32_t q_head = 0, q_cnt = 0;  // warp-uniform: every lane derives them from the same ballots
        // test the queued point of this lane against its candidate boxes; accumulate the (usually single) containing box once
        // test the queued point of this lane against its candidate boxes; accumulate the (usually single) containing box once
        auto drain_queue = [&](uint32_t n_take) {
            const bool act = (uint32_t)lane < n_take;
            const float4 e = queue[(q_head + lane) & 63u];
            __syncwarp();  // every lane has read its slot before any lane can enqueue over it again
            if (act) {
                uint32_t ids = __float_as_uint(e.w);
                const float es2 = __fadd_rn(__fmul_rn(e.x, e.x), __fmul_rn(e.y, e.y));
                int hit = -1;
                do {
                    const int b = (int)(ids & 0xffu);
                    if (box_contains(boxp, b, e.x, e.y, e.z)) {
                        if (hit >= 0) box_accumulate(A, boxacc, b, e.x, e.y, e.z, es2); else hit = b;
                    }
                    ids = (ids >> 8) | 0xff000000u;
                } while ((ids & 0xffu) != 0xffu);
                if (hit >= 0 && !(A.debug_skip & 8u)) box_accumulate(A, boxacc, hit, e.x, e.y, e.z, es2);
            }
            q_head = (q_head + n_take) & 63u;
            q_cnt -= n_take;
        };
        double M[MSMEM ? 1 : 12];
        const double* Ms = nullptr;
        for (;;) {
            issue_next();  // refill the slot consumed in the previous iteration (S - 1 tiles ahead)
            bool done = false;
            while (c_first >= c_cnt) {
                if (++c_si >= n_sw) { done = true; break; }
                c_cnt = sweep_cnt(c_si); c_first = (uint32_t)warp * TP;
                if (c_first < c_cnt) {
                    if (MSMEM) Ms = (c_si < kMaxSweepsSmem) ? (misc->pose + c_si * 12) : (A.in.sweep_pose + (size_t)(sw0 + c_si) * 12);
                    else ld_pose(A.in.sweep_pose + (size_t)(sw0 + c_si) * 12, M);
                }
            }
            if (done) break;
            const uint32_t npts = min((uint32_t)TP, c_cnt - c_first);
            const int stage = (int)(wk % S);
            mbar_wait_parity(&full[stage], (wk / S) & 1u);
            const float* tp = ring + (size_t)stage * (C::kTileBytes / 4) + lane * 5;

            // ---- phase A: branch-free over the lane's PPT points so their dependency chains interleave
            float xr[PPT], yr[PPT], zr[PPT], s2[PPT], inten[PPT];
            int ix[PPT], iy[PPT];
            uint2 ce[PPT];
            bool keep[PPT], alive[PPT];
            double xd[PPT], yd[PPT], zd[PPT];
#pragma unroll
            for (int u = 0; u < PPT; ++u) {
                const bool valid = (uint32_t)lane + (uint32_t)u * 32u < npts;
                const float x = tp[u * 160 + 0], y = tp[u * 160 + 1], z = tp[u * 160 + 2];
                inten[u] = tp[u * 160 + 3];
                // A.1 remove_close (square, sweep's own sensor frame)
                alive[u] = valid && !(fabsf(x) < P.remove_close_radius && fabsf(y) < P.remove_close_radius);
                xd[u] = (double)x; yd[u] = (double)y; zd[u] = (double)z;
                c_close += alive[u] ? 1u : 0u;
            }
            __syncwarp();  // every lane has read its rows: the slot may be refilled at the top of the next iteration
            // A.1 f64 matrix x f32 point -> f32, one matrix row at a time (keeps few pose values live when they come from smem)
            if (MSMEM) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const double2 m01 = *reinterpret_cast<const double2*>(Ms + r * 4), m23 = *reinterpret_cast<const double2*>(Ms + r * 4 + 2);
#pragma unroll
                    for (int u = 0; u < PPT; ++u) {
                        const float v = (float)__fma_rn(m01.x, xd[u], __fma_rn(m01.y, yd[u], __fma_rn(m23.x, zd[u], m23.y)));
                        if (r == 0) xr[u] = v; else if (r == 1) yr[u] = v; else zr[u] = v;
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < PPT; ++u) {
                    xr[u] = (float)__fma_rn(M[0], xd[u], __fma_rn(M[1], yd[u], __fma_rn(M[2], zd[u], M[3])));
                    yr[u] = (float)__fma_rn(M[MSMEM ? 0 : 4], xd[u], __fma_rn(M[MSMEM ? 0 : 5], yd[u], __fma_rn(M[MSMEM ? 0 : 6], zd[u], M[MSMEM ? 0 : 7])));
                    zr[u] = (float)__fma_rn(M[MSMEM ? 0 : 8], xd[u], __fma_rn(M[MSMEM ? 0 : 9], yd[u], __fma_rn(M[MSMEM ? 0 : 10], zd[u], M[MSMEM ? 0 : 11])));
                }
            }
#pragma unroll
            for (int u = 0; u < PPT; ++u) {
                // lidar_agent.py:106-110, sqrt-free (thresholds on s are exact, geometry.sqrt_thresholds)
                s2[u] = __fadd_rn(__fmul_rn(xr[u], xr[u]), __fmul_rn(yr[u], yr[u]));
                keep[u] = alive[u] && (s2[u] >= P.s_lo) && (s2[u] <= P.s_hi) && (zr[u] < P.z_max) && (zr[u] > P.z_min);
                // BEV cell, lidar_agent.py:547-552 (garbage for dropped points is clamped and never used)
                ix[u] = bev_cell<FASTDIV>(xr[u], P.bev_range, A.two_r, A.rcp_two_r, A.resf, res_m1);
                iy[u] = bev_cell<FASTDIV>(yr[u], P.bev_range, A.two_r, A.rcp_two_r, A.resf, res_m1);
                ce[u] = cull[(iy[u] >> L.cull_shift) * L.cull_dim + (ix[u] >> L.cull_shift)];
            }
            // ---- phase B: data-dependent work per kept point
            uint32_t cand[PPT];
#pragma unroll
            for (int u = 0; u < PPT; ++u) {
                cand[u] = kCullEmpty;
                if (!keep[u]) continue;
                if (FOV) {
                    uint32_t in_bits = ce[u].y & 0xffu;
                    uint32_t st = ce[u].y >> 8;
                    while (st) {  // exact wedge test only where the cell straddles a wedge edge
                        const int c = __ffs((int)st) - 1;
                        st &= st - 1;
                        in_bits |= in_wedge(misc->wedge[c], xr[u], yr[u]) ? (1u << c) : 0u;
                    }
                    // spread 8 bits into 8 byte counters (no carries: the multiplier's partial products do not overlap)
                    cam_lo += ((in_bits & 0xfu) * 0x00204081u) & 0x01010101u;
                    cam_hi += ((in_bits >> 4) * 0x00204081u) & 0x01010101u;
                    if (P.fov_keep_mask != 0u && (in_bits & P.fov_keep_mask) == 0u) continue;
                }
                ++c_kept;
                c_ground += (zr[u] < P.ground_z) ? 1u : 0u;  // lidar_agent.py:128
                // Q8 intensity, clamp [0, 65535]; NaN -> 0
                const float qf = fminf(fmaxf(__fmul_rn(inten[u], A.iscale), 0.0f), 65535.0f);
                const uint32_t q = (uint32_t)__float2int_rn(qf);
                const uint32_t wx = (uint32_t)(ix[u] - L.win_lo), wy = (uint32_t)(iy[u] - L.win_lo);
                const uint32_t cell = (uint32_t)iy[u] * (uint32_t)res + (uint32_t)ix[u];
                if (wx < (uint32_t)L.win_w && wy < (uint32_t)L.win_w) {
                    uint2* wc = window + wy * (uint32_t)L.win_w + wx;
                    if (!(A.debug_skip & 4u)) {
                        atomicAdd(&wc->x, 1u);
                        atomicAdd(&wc->y, q);
                    }
                } else if (!(A.debug_skip & 1u)) {
                    atomicAdd(reinterpret_cast<unsigned long long*>(g_ci) + cell, 1ull | ((unsigned long long)q << 32));
                }
                if (zr[u] > 0.0f && !(A.debug_skip & 1u)) atomicMax(reinterpret_cast<int*>(g_h) + cell, __float_as_int(zr[u]));  // :560, 0-initialised max
                if (C::kQueue) { cand[u] = (A.debug_skip & 2u) ? kCullEmpty : ce[u].x; continue; }  // enqueued after the loop
                // A.2 oriented-box membership for the candidate boxes of this cull cell
                uint32_t ids = ce[u].x;
                if (ids == kCullEmpty || (A.debug_skip & 2u)) continue;
                // The divergent candidate loop only tests; the accumulator update of the (usually single) containing box runs
                // once per point after the loop.  A second containing box (overlapping annotations) updates inside the loop.
                int hit = -1;
                if (ids == kCullAll) {  // crowded cell (more than four boxes): test every box
                    for (int b = 0; b < n_boxes; ++b)
                        if (box_contains(boxp, b, xr[u], yr[u], zr[u])) {
                            if (hit >= 0) box_accumulate(A, boxacc, b, xr[u], yr[u], zr[u], s2[u]); else hit = b;
                        }
                } else {
                    do {
                        const int b = (int)(ids & 0xffu);
                        if (box_contains(boxp, b, xr[u], yr[u], zr[u])) {
                            if (hit >= 0) box_accumulate(A, boxacc, b, xr[u], yr[u], zr[u], s2[u]); else hit = b;
                        }
                        ids = (ids >> 8) | 0xff000000u;
                    } while ((ids & 0xffu) != 0xffu);
                }
                if (hit >= 0 && !(A.debug_skip & 8u)) box_accumulate(A, boxacc, hit, xr[u], yr[u], zr[u], s2[u]);
            }
            if (C::kQueue) {
                // ---- phase C: points that have candidate boxes go to this warp's queue; whenever 32 are pending every lane tests
                // one of them (dense), instead of a handful of lanes looping while the rest of the warp idles
#pragma unroll
                for (int u = 0; u < PPT; ++u) {
                    if (cand[u] == kCullAll) {  // crowded cell (more than four boxes): rare, test every box in place
                        int hit = -1;
                        for (int b = 0; b < n_boxes; ++b)
                            if (box_contains(boxp, b, xr[u], yr[u], zr[u])) {
                                if (hit >= 0) box_accumulate(A, boxacc, b, xr[u], yr[u], zr[u], s2[u]); else hit = b;
                            }
                        if (hit >= 0 && !(A.debug_skip & 8u)) box_accumulate(A, boxacc, hit, xr[u], yr[u], zr[u], s2[u]);
                        cand[u] = kCullEmpty;
                    }
                    const bool has = cand[u] != kCullEmpty;
                    const uint32_t m = __ballot_sync(0xffffffffu, has);
                    if (m == 0u) continue;
                    const uint32_t add = __popc(m);
                    if (q_cnt + add > 64u) drain_queue(32u);
                    if (has) queue[(q_head + q_cnt + __popc(m & ((1u << lane) - 1u))) & 63u] = make_float4(xr[u], yr[u], zr[u], __uint_as_float(cand[u]));
                    q_cnt += add;
                    __syncwarp();
                    if (q_cnt >= 32u) drain_queue(32u);
                }
            }
            ++wk;
            c_first += W * TP;
            if (FOV) {
                cam_pts += PPT;
                if (cam_pts > 255u - PPT) {  // spill the byte counters before any of them can wrap
#pragma unroll
                    for (int c = 0; c < MSC_MAX_CAMS; ++c) {
                        const uint32_t v = ((c < 4 ? cam_lo : cam_hi) >> ((c & 3) * 8)) & 0xffu;
                        if (v) atomicAdd(&misc->stats[5 + c], v);
                    }
                    cam_lo = cam_hi = cam_pts = 0;
                }
            }
        }

        if (C::kQueue) {
            while (q_cnt > 0u) drain_queue(min(q_cnt, 32u));
        }

        // ------------------------------------------------------------ epilogue
        {
            uint32_t v[3 + MSC_MAX_CAMS];
            v[0] = c_close; v[1] = c_kept; v[2] = c_ground;
#pragma unroll
            for (int c = 0; c < MSC_MAX_CAMS; ++c) v[3 + c] = FOV ? (((c < 4 ? cam_lo : cam_hi) >> ((c & 3) * 8)) & 0xffu) : 0u;
#pragma unroll
            for (int i = 0; i < 3 + (FOV ? MSC_MAX_CAMS : 0); ++i) {
                const uint32_t r = __reduce_add_sync(0xffffffffu, v[i]);
                if (lane == 0 && r) atomicAdd(&misc->stats[i < 3 ? 1 + i : 2 + i], r);
            }
        }
        __syncthreads();  // every tile of the sample is accumulated
        {
            // window flush: coalesced 16-byte stores of (count, isum) pairs, two cells per store
            const int half_w = L.win_w >> 1;  // win_w and win_lo are even -> 16-byte aligned rows
            uint32_t flags = 0;
            for (int i = tid; i < L.win_w * half_w; i += NT) {
                const int wy = i / half_w, wx2 = i - wy * half_w;
                const uint4 v = reinterpret_cast<const uint4*>(window)[wy * half_w + wx2];
                const size_t cell = (size_t)(wy + L.win_lo) * (size_t)res + (size_t)(wx2 * 2 + L.win_lo);
                *reinterpret_cast<uint4*>(g_ci + cell * 2) = v;
                flags |= (v.x >= 65536u || v.z >= 65536u) ? 1u : 0u;
            }
            // per-box results
            for (int b = tid; b < n_boxes; b += NT) {
                const uint32_t* acc = boxacc + b * kAccWords;
                const uint32_t cnt = acc[0];
                const size_t o = (size_t)(bx0 + b);
                A.out.box_count[o] = cnt;
                if (cnt >= (1u << 20)) flags |= 2u;  // 12-bit limb sums may have wrapped
                if (cnt == 0) {
                    A.out.box_nearest[o] = INFINITY;
                    A.out.box_centroid[o * 3 + 0] = 0.0f; A.out.box_centroid[o * 3 + 1] = 0.0f; A.out.box_centroid[o * 3 + 2] = 0.0f;
                } else {
                    A.out.box_nearest[o] = __fsqrt_rn(__uint_as_float(acc[1]));
                    const double den = (double)cnt * (double)A.cscale;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const unsigned long long biased = (unsigned long long)acc[2 + 2 * k] + ((unsigned long long)acc[3 + 2 * k] << 12);
                        const long long sum = (long long)biased - (long long)cnt * (long long)A.centroid_bias;
                        A.out.box_centroid[o * 3 + k] = (float)((double)sum / den);
                    }
                }
            }
            if (flags) atomicOr(&misc->stats[13], flags);
        }
        __syncthreads();
        if (tid < MSC_STATS_STRIDE) {
            uint32_t v = misc->stats[tid];
            if (tid == 0) {
                v = 0;
                for (int s = sw0; s < sw1; ++s) v += A.in.sweep_count[s];
            }
            if (tid == 4) v = misc->stats[2] - misc->stats[3];  // n_object = n_kept - n_ground
            if (tid == 13 && box_overflow) v |= 0x80000000u;
            A.out.stats[(size_t)sample * MSC_STATS_STRIDE + tid] = v;
        }
        // (the __syncthreads after the next sample fetch orders these reads before the smem is re-zeroed)
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static int g_opt_fov = 1;
static int g_opt_window = 0;       // 0 = auto (largest that fits)
static int g_opt_cull_shift = -1;  // -1 = auto (cull cell ~ 2 m)
static int g_opt_fastdiv = 1;      // allow the Markstein division for whitelisted divisors
static int g_opt_debug_skip = 0;
static int g_opt_config = 7;  // launch shape: 7 = second-generation kernel (fused_stream.cu), 1024 threads x 2 points per lane (default);
                              // 8 = the same kernel with 512 threads x 4 points per lane; first generation (this file): 6 = 1024 x 2, pose
                              // rows in smem, candidate queue (its final shape); any other value = 512 x 2, four ring stages, pose in registers
static int g_opt_time_kernel = 0;  // bracket the streaming kernel with CUDA events (msc_fused_kernel_times)
static int g_last_launches = 0;    // kernels launched by the most recent msc_fused_evidence_batch call
constexpr int kTimeRing = 64;
static cudaEvent_t g_ev0[kTimeRing], g_ev1[kTimeRing];
static bool g_ev_made = false;
static long long g_ev_count = 0;  // calls timed so far
static int time_begin(cudaStream_t stream) {
    if (!g_opt_time_kernel) return MSC_OK;
    if (!g_ev_made) {
        for (int i = 0; i < kTimeRing; ++i) { MSC_CUDA(cudaEventCreate(&g_ev0[i])); MSC_CUDA(cudaEventCreate(&g_ev1[i])); }
        g_ev_made = true;
    }
    MSC_CUDA(cudaEventRecord(g_ev0[g_ev_count % kTimeRing], stream));
    return MSC_OK;
}
static int time_end(cudaStream_t stream) {
    if (!g_opt_time_kernel) return MSC_OK;
    MSC_CUDA(cudaEventRecord(g_ev1[g_ev_count % kTimeRing], stream));
    ++g_ev_count;
    return MSC_OK;
}
static int g_last_window = 0, g_last_smem = 0, g_last_fastdiv = 0, g_last_tile_pts = 0, g_last_stages = 0, g_last_threads = 0;

// divisors 2*bev_range for which tools/markstein_check.c has been run over the full float range
static bool fastdiv_verified(float two_r) {
    const float ok[] = {100.0f, 102.4f, 120.0f, 150.0f, 160.0f, 200.0f};
    for (float v : ok)
        if (two_r == v) return true;
    int e = 0;
    return frexpf(two_r, &e) == 0.5f;  // powers of two divide exactly either way
}

static void cull_geometry(const msc_params& P, int* shift, int* dim) {
    const float cell_m = 2.0f * P.bev_range / (float)P.bev_res;
    int sh = 0;
    if (g_opt_cull_shift >= 0) sh = g_opt_cull_shift;
    else while ((float)(1 << (sh + 1)) * cell_m <= 2.0f + 1e-6f && sh < 10) ++sh;
    *shift = sh;
    *dim = ((P.bev_res - 1) >> sh) + 1;
}

static TableLayout table_layout(const msc_params& P, int n_samples, int n_boxes) {
    int shift, dim;
    cull_geometry(P, &shift, &dim);
    auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
    TableLayout T;
    size_t off = 0;
    T.counter_off = off; off = align(off + 256);
    T.boxprep_off = off; off = align(off + (size_t)(n_boxes > 0 ? n_boxes : 1) * kBoxStride * 4);
    T.wedge_off = off; off = align(off + (size_t)(n_samples > 0 ? n_samples : 1) * MSC_MAX_CAMS * 6 * 4);
    T.fovcls_off = off; off = align(off + (size_t)(n_samples > 0 ? n_samples : 1) * dim * dim * 2);
    T.edgecls_off = off; off = align(off + (size_t)(n_samples > 0 ? n_samples : 1) * dim * dim * 4);
    T.innercls_off = off; off = align(off + (size_t)(n_samples > 0 ? n_samples : 1) * kInnerMax * kInnerMax * 4);
    T.cullids_off = off; off = align(off + (size_t)(n_samples > 0 ? n_samples : 1) * dim * dim * 4);
    T.total = off;
    return T;
}

static int compute_layout(const msc_params& P, int max_boxes_in_batch, int smem_limit, int ring_bytes, int queue_bytes, int misc_bytes, int inner_dim,
                          FusedLayout* L) {
    const int cap = max_boxes_in_batch < 1 ? 1 : max_boxes_in_batch;
    cull_geometry(P, &L->cull_shift, &L->cull_dim);
    L->max_boxes = cap;
    int off = 0;
    L->tiles_off = off; off += ring_bytes; off = (off + 127) & ~127;
    L->queue_off = off; off += queue_bytes; off = (off + 127) & ~127;
    L->cull_off = off; off += L->cull_dim * L->cull_dim * 8; off = (off + 127) & ~127;
    L->boxp_off = off; off += cap * kBoxStride * 4;
    L->boxacc_off = off; off += cap * kAccWords * 4; off = (off + 127) & ~127;
    L->misc_off = off; off += (misc_bytes + 127) & ~127;
    L->inner_dim = inner_dim; L->inner_lo = (P.bev_res - inner_dim) / 2;
    L->inner_off = off; off += (inner_dim * inner_dim * 4 + 127) & ~127;
    L->window_off = off;
    const int avail = smem_limit - off;
    if (avail < 0) return -1;
    int w = 0;
    while ((w + 2) * (w + 2) * 8 <= avail && (w + 2) <= P.bev_res) w += 2;
    if (g_opt_window > 0 && g_opt_window < w) w = g_opt_window & ~1;
    if (((P.bev_res - w) / 2) & 1) w -= 2;  // keep win_lo even so flush rows stay 16-byte aligned
    if (w < 0) w = 0;
    L->win_w = w;
    L->win_lo = (P.bev_res - w) / 2;
    L->total_bytes = off + w * w * 8;
    return 0;
}

template <class C, bool FOV, bool FASTDIV>
static int launch_fused(const FusedArgs& args, const TableLayout& T, unsigned char* ws, int grid, cudaStream_t stream) {
    auto kern = fused_evidence_kernel<C, FOV, FASTDIV>;
    MSC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, args.L.total_bytes));
    kern<<<grid, C::kThreads, args.L.total_bytes, stream>>>(args, T, ws);
    MSC_CUDA(cudaGetLastError());
    return MSC_OK;
}
// tables (prepared boxes, projection, wedges, wedge classes -> workspace), launched before either streaming kernel
// one side stream per device: the class kernel runs beside the cull-id kernel (both only need the table kernel's output)
struct SideStream { bool made = false; cudaStream_t s; cudaEvent_t fork, join; };
static SideStream g_side[64];
static int side_stream(SideStream** out) {
    int dev = 0;
    MSC_CUDA(cudaGetDevice(&dev));
    SideStream& S = g_side[dev & 63];
    if (!S.made) {
        MSC_CUDA(cudaStreamCreateWithFlags(&S.s, cudaStreamNonBlocking));
        MSC_CUDA(cudaEventCreateWithFlags(&S.fork, cudaEventDisableTiming));
        MSC_CUDA(cudaEventCreateWithFlags(&S.join, cudaEventDisableTiming));
        S.made = true;
    }
    *out = &S;
    return MSC_OK;
}

// `side` != nullptr: the class kernel goes to the side stream and the caller joins it (cudaStreamWaitEvent(stream, side->join))
// before its streaming kernel
static int launch_tables(const FusedArgs& args, const TableLayout& T, unsigned char* ws, int n_boxes_total, bool fov, bool per_edge,
                         cudaStream_t stream, SideStream* side = nullptr) {
    const int ncc = args.L.cull_dim * args.L.cull_dim;
    const int cams = args.P.n_cams > 0 ? args.P.n_cams : 1;
    long long work = (long long)n_boxes_total * cams;
    if ((long long)args.in.n_samples * cams > work) work = (long long)args.in.n_samples * cams;
    g_last_launches = 0;
    if (work > 0) {
        fused_tables_kernel<<<(unsigned)((work + 255) / 256), 256, 0, stream>>>(args, T, n_boxes_total, ws);
        MSC_CUDA(cudaGetLastError());
        ++g_last_launches;
    }
    if (fov) {
        const long long cells = (long long)args.in.n_samples * (ncc + (per_edge ? args.L.inner_dim * args.L.inner_dim : 0));
        cudaStream_t cs = stream;
        if (side) {
            MSC_CUDA(cudaEventRecord(side->fork, stream));
            MSC_CUDA(cudaStreamWaitEvent(side->s, side->fork, 0));
            cs = side->s;
        }
        fused_fovcls_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, cs>>>(args, T, ws, per_edge);
        MSC_CUDA(cudaGetLastError());
        if (side) MSC_CUDA(cudaEventRecord(side->join, side->s));
        ++g_last_launches;
    }
    return MSC_OK;
}

template <class C>
static int dispatch_fused(FusedArgs& args, const TableLayout& T, unsigned char* ws, int n_boxes_total, int smem_optin, int grid, bool fov,
                          bool fast, cudaStream_t stream) {
    if (compute_layout(args.P, args.in.max_boxes_per_sample, smem_optin, C::kRingBytes, C::kQueueBytes, (int)sizeof(Misc), 0, &args.L) != 0) {
        set_error("shared-memory layout does not fit (%d bytes available)", smem_optin);
        return MSC_ERR_UNSUPPORTED;
    }
    g_last_window = args.L.win_w;
    g_last_smem = args.L.total_bytes;
    g_last_tile_pts = C::kTilePts; g_last_stages = C::kStages; g_last_threads = C::kThreads;
    int rc = launch_tables(args, T, ws, n_boxes_total, fov, false, stream);
    if (rc != MSC_OK) return rc;
    if ((rc = time_begin(stream)) != MSC_OK) return rc;
    if (fov) rc = fast ? launch_fused<C, true, true>(args, T, ws, grid, stream) : launch_fused<C, true, false>(args, T, ws, grid, stream);
    else rc = fast ? launch_fused<C, false, true>(args, T, ws, grid, stream) : launch_fused<C, false, false>(args, T, ws, grid, stream);
    if (rc != MSC_OK) return rc;
    ++g_last_launches;
    return time_end(stream);
}

// configs 7-8: the second-generation streaming kernel (fused_stream.cu), launch shapes 0-1
static int dispatch_stream(int shape, FusedArgs& args, const TableLayout& T, unsigned char* ws, int n_boxes_total, int smem_optin, int grid,
                           bool fov, bool fast, cudaStream_t stream) {
    int threads = 0, tile_pts = 0, ring = 0, queue = 0;
    stream_shape_info(shape, &threads, &tile_pts, &ring, &queue);
    // fine edge classes for the kInnerMax x kInnerMax BEV cells around the sensor, where several image-column rays cross a 2 m cull cell
    int inner = fov ? (args.P.bev_res < kInnerMax ? args.P.bev_res : kInnerMax) : 0;
    inner &= ~1;
    if (compute_layout(args.P, args.in.max_boxes_per_sample, smem_optin, ring, queue, stream_misc_bytes(), inner, &args.L) != 0) {
        set_error("shared-memory layout does not fit (%d bytes available)", smem_optin);
        return MSC_ERR_UNSUPPORTED;
    }
    g_last_window = args.L.win_w;
    g_last_smem = args.L.total_bytes;
    g_last_tile_pts = tile_pts; g_last_stages = 2; g_last_threads = threads;
    SideStream* side = nullptr;
    int rc = side_stream(&side);
    if (rc != MSC_OK) return rc;
    if ((rc = launch_tables(args, T, ws, n_boxes_total, fov, true, stream, side)) != MSC_OK) return rc;
    const size_t ncc = (size_t)args.L.cull_dim * args.L.cull_dim;
    MSC_CUDA(cudaMemsetAsync(ws + T.cullids_off, 0xff, (size_t)args.in.n_samples * ncc * 4, stream));  // kCullEmpty
    if (n_boxes_total > 0 && args.L.max_boxes > 0) {
        const dim3 cgrid((unsigned)args.in.n_samples, (unsigned)((args.L.max_boxes + 3) / 4));  // a warp per box, one grid column per sample
        fused_cullids_kernel<<<cgrid, 128, 0, stream>>>(args, T, ws);
        MSC_CUDA(cudaGetLastError());
        ++g_last_launches;
    }
    if (fov) MSC_CUDA(cudaStreamWaitEvent(stream, side->join, 0));  // the class tables are ready
    if ((rc = time_begin(stream)) != MSC_OK) return rc;
    if ((rc = launch_stream_kernel(shape, args, T, ws, grid, fov, fast, stream)) != MSC_OK) return rc;
    ++g_last_launches;
    return time_end(stream);
}

}  // namespace msc

extern "C" {

size_t msc_fused_workspace_bytes(const msc_params* params, int32_t n_samples, int32_t n_boxes) {
    if (!params) return 0;
    return msc::table_layout(*params, n_samples, n_boxes).total;
}

int msc_fused_set_option(const char* key, int32_t value) {
    if (!key) return MSC_ERR_BAD_ARGUMENT;
    if (!strcmp(key, "fov")) { msc::g_opt_fov = value ? 1 : 0; return MSC_OK; }
    if (!strcmp(key, "window")) { msc::g_opt_window = value; return MSC_OK; }
    if (!strcmp(key, "cull_shift")) { msc::g_opt_cull_shift = value; return MSC_OK; }
    if (!strcmp(key, "fastdiv")) { msc::g_opt_fastdiv = value ? 1 : 0; return MSC_OK; }
    if (!strcmp(key, "config")) { msc::g_opt_config = value; return MSC_OK; }
    if (!strcmp(key, "debug_skip")) { msc::g_opt_debug_skip = value; return MSC_OK; }
    if (!strcmp(key, "time_kernel")) { msc::g_opt_time_kernel = value ? 1 : 0; return MSC_OK; }
    msc::set_error("unknown option %s", key);
    return MSC_ERR_BAD_ARGUMENT;
}

int msc_fused_get_option(const char* key, int32_t* value) {
    if (!key || !value) return MSC_ERR_BAD_ARGUMENT;
    if (!strcmp(key, "fov")) { *value = msc::g_opt_fov; return MSC_OK; }
    if (!strcmp(key, "window")) { *value = msc::g_opt_window; return MSC_OK; }
    if (!strcmp(key, "cull_shift")) { *value = msc::g_opt_cull_shift; return MSC_OK; }
    if (!strcmp(key, "fastdiv")) { *value = msc::g_opt_fastdiv; return MSC_OK; }
    if (!strcmp(key, "config")) { *value = msc::g_opt_config; return MSC_OK; }
    if (!strcmp(key, "last_window")) { *value = msc::g_last_window; return MSC_OK; }
    if (!strcmp(key, "last_smem")) { *value = msc::g_last_smem; return MSC_OK; }
    if (!strcmp(key, "last_fastdiv")) { *value = msc::g_last_fastdiv; return MSC_OK; }
    if (!strcmp(key, "tile_pts")) { *value = msc::g_last_tile_pts; return MSC_OK; }
    if (!strcmp(key, "stages")) { *value = msc::g_last_stages; return MSC_OK; }
    if (!strcmp(key, "threads")) { *value = msc::g_last_threads; return MSC_OK; }
    if (!strcmp(key, "last_launches")) { *value = msc::g_last_launches; return MSC_OK; }
    if (!strcmp(key, "time_kernel")) { *value = msc::g_opt_time_kernel; return MSC_OK; }
    msc::set_error("unknown option %s", key);
    return MSC_ERR_BAD_ARGUMENT;
}

int msc_fused_kernel_times(float* out_ms_host, int32_t n) {
    using namespace msc;
    MSC_REQUIRE(out_ms_host && n >= 0, "bad argument");
    const long long have = g_ev_count < kTimeRing ? g_ev_count : kTimeRing;
    const int take = (int)(n < have ? n : have);
    for (int i = 0; i < take; ++i) {
        const long long k = g_ev_count - take + i;
        MSC_CUDA(cudaEventSynchronize(g_ev1[k % kTimeRing]));
        MSC_CUDA(cudaEventElapsedTime(out_ms_host + i, g_ev0[k % kTimeRing], g_ev1[k % kTimeRing]));
    }
    return take;
}

int msc_fused_evidence_batch(const msc_params* params, const msc_batch_in* in, const msc_batch_out* out, void* workspace,
                             size_t workspace_bytes, void* stream_v) {
    using namespace msc;
    MSC_REQUIRE(params && in && out && workspace, "null argument");
    MSC_REQUIRE(in->n_samples >= 0 && in->n_boxes >= 0, "negative counts");
    MSC_REQUIRE(params->n_cams >= 0 && params->n_cams <= MSC_MAX_CAMS, "n_cams out of range");
    MSC_REQUIRE(params->bev_res > 0 && params->bev_res <= 4096 && (params->bev_res & 1) == 0, "bev_res must be even and <= 4096");
    MSC_REQUIRE(params->centroid_shift >= 0 && params->centroid_shift <= 17, "centroid_shift out of range (two 12-bit limbs hold 24 bits)");
    MSC_REQUIRE(params->intensity_shift >= 0 && params->intensity_shift <= 8, "intensity_shift out of range");
    // the biased fixed-point coordinate must fit 24 bits: |c| * 2^shift < 2^(shift + 6)  <=>  |c| < 64 m
    MSC_REQUIRE(params->range_max < 64.0f && params->z_max < 64.0f && params->z_min > -64.0f, "range_max / z limits must be below 64 m");
    MSC_REQUIRE(in->max_boxes_per_sample >= 0 && in->max_boxes_per_sample <= MSC_MAX_BOXES_FUSED, "more than %d boxes in one sample",
                MSC_MAX_BOXES_FUSED);
    MSC_REQUIRE((((uintptr_t)in->points) & 15) == 0, "points must be 16-byte aligned");
    MSC_REQUIRE((((uintptr_t)workspace) & 255) == 0, "workspace must be 256-byte aligned");
    const TableLayout T = table_layout(*params, in->n_samples, in->n_boxes);
    MSC_REQUIRE(workspace_bytes >= T.total, "workspace too small: need %zu bytes", T.total);
    cudaStream_t stream = (cudaStream_t)stream_v;
    if (in->n_samples == 0) return MSC_OK;
    int dev = 0, sms = 0, smem_optin = 0;
    MSC_CUDA(cudaGetDevice(&dev));
    MSC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MSC_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    FusedArgs args;
    args.P = *params;
    args.in = *in;
    args.out = *out;
    args.two_r = 2.0f * params->bev_range;
    args.resf = (float)params->bev_res;
    args.rcp_two_r = 1.0f / args.two_r;
    args.cscale = (float)(1 << params->centroid_shift);
    args.iscale = (float)(1 << params->intensity_shift);
    args.centroid_bias = 1 << (params->centroid_shift + 6);
    args.debug_skip = (uint32_t)g_opt_debug_skip;
    const bool fov = g_opt_fov != 0 && params->n_cams > 0;
    const bool fast = g_opt_fastdiv != 0 && fastdiv_verified(args.two_r);
    g_last_fastdiv = fast ? 1 : 0;
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    MSC_CUDA(cudaMemsetAsync(ws + T.counter_off, 0, 256, stream));
    const int grid = in->n_samples < sms ? in->n_samples : sms;
    switch (g_opt_config) {
        case 7: case 8: return dispatch_stream(g_opt_config - 7, args, T, ws, in->n_boxes, smem_optin, grid, fov, fast, stream);
        case 6: return dispatch_fused<Cfg<1024, 2, 2, true, 1>>(args, T, ws, in->n_boxes, smem_optin, grid, fov, fast, stream);
        default: return dispatch_fused<Cfg<512, 2, 4>>(args, T, ws, in->n_boxes, smem_optin, grid, fov, fast, stream);
    }
}

}  // extern "C"
