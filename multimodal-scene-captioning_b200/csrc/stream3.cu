// stream3.cu -- third generation of the streaming kernel (launch shape "config" 9, the default).
//
// Same arithmetic per point as fused_stream.cu (SURVEY.md App. A + lidar_agent.py:103-132, :547-560).  What changed, and why
// (profiles/r1b_stream_evidence_ncu_full.txt: issue-bound at 530 warp instructions per 64-point tile, 42 % of the shared-memory
// wavefronts bank-conflict replays, ~100 of the 530 instructions spent on per-camera FOV counts):
//   * FOV counts are DERIVED per BEV cell in the epilogue: a cell every camera has decided (81 % of the kept points) contributes
//     count[cell] x in-bits; only the points of undecided cells evaluate a cross product.  Which edge a point has to test arrives
//     for free: the top five bits of its window count word hold the cell's "edge code" (preset in the prologue), and the shared
//     atomic that counts the point returns them.  One branch-free test per point, a cold path for the rare cell two rays cross.
//   * n_kept is the sum of the cell counts (epilogue), rows past the end of a partial tile are NaN in registers (no per-point
//     validity compare), the Q8 intensity is a clamp + FFMA magic-number rounding (no conversion-pipe instruction).
//   * the window keeps count and intensity sum in two arrays (all 32 banks per atomic instead of 16), dropped points add into
//     per-lane sink words in front of each array.
//   * one sample can be SPLIT over several CTAs (split > 1: a single keyframe, or a shard smaller than the SM count): every part
//     streams a share of the tiles, merges its integer accumulators into global memory with reductions, and the part that takes the
//     last ticket finalises the sample.  Integer accumulators make the result independent of the split.
#include "fused_common.cuh"

namespace msc {

constexpr int kS3Threads = 1024, kS3Warps = 32, kS3Ppt = 2, kS3TilePts = 64, kS3TileFloats = 320, kS3TileBytes = 1280;
constexpr int kS3RingBytes = kS3Warps * 2 * kS3TileBytes;
constexpr int kS3QueueBytes = kS3Warps * 64 * 16;
constexpr int kS3PoseSmem = 12;  // sweeps whose transforms are staged per sample; later ones use a per-warp slot
constexpr uint32_t kCodeShift = 27, kCountMask = (1u << kCodeShift) - 1u, kCodeMulti = 31u;

struct alignas(16) S3Edge {  // one entry per edge code: the exact test and the byte-counter increments of its camera
    float ax, ay, a, b;      // cr = fma(a, s, -(b * t)); (s, t) = (qy, qx) for a right edge, (qx, qy) for a left edge
    uint32_t inc_lo, inc_hi, pad0, pad1;
};

struct S3Misc {  // small per-CTA state at misc_off
    uint64_t full_bar[kS3Warps * 2];  // [warp][slot]: TMA bytes landed in that warp's ring slot
    S3Edge edge[32];                  // [code]: 0 and 17..31 are pads whose test fails (a = NaN); 1 + c right edge, 9 + c left edge of camera c
    uint32_t stats[MSC_STATS_STRIDE];
    uint32_t sweep_start[kS3PoseSmem], sweep_count[kS3PoseSmem];
    int32_t sample, part, ticket, pad_;
    unsigned long long* ci64;         // this sample's (count, isum) layer as 64-bit cells and its max-height layer, read back by
    int* h32;                         // the few lanes per warp that update cells outside the window / the height layer
    double pose[kS3PoseSmem * 12];    // this sample's 3x4 sweep transforms
    double wpose[kS3Warps * 12];      // per-warp slot for sweeps beyond kS3PoseSmem
    float wq[MSC_MAX_CAMS * 6];       // this sample's camera wedges (fused_tables_kernel), source of the per-cell edge classes
};

int stream3_misc_bytes() { return (int)sizeof(S3Misc); }
int stream3_ring_bytes() { return kS3RingBytes; }
int stream3_queue_bytes() { return kS3QueueBytes; }

// ---- shared-state-space accesses through 32-bit addresses (no generic-address arithmetic in the loop)
__device__ __forceinline__ void s3_red_add(uint32_t saddr, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t s3_atom_add(uint32_t saddr, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(saddr), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ float4 s3_lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint2 s3_lds64(uint32_t saddr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void s3_lds_f64x2(uint32_t saddr, double& a, double& b) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(saddr));
}

// edge code of a cell from its class word (bits 0-7 in-bits, 8-15 right edge undecided, 16-23 left edge undecided):
// 0 = every camera decided, 1 + e = exactly edge e (0-7 right, 8-15 left) undecided, 31 = several
__device__ __forceinline__ uint32_t s3_code_of(uint32_t cls) {
    const uint32_t und = (cls >> 8) & 0xffffu;
    if (und == 0u) return 0u;
    if (und & (und - 1u)) return kCodeMulti;
    return (uint32_t)__ffs((int)und);
}
// cameras whose wedge contains every point of the cell
__device__ __forceinline__ uint32_t s3_decided_in(uint32_t cls) { return cls & ~(cls >> 8) & ~(cls >> 16) & 0xffu; }

template <bool FOV, bool FASTDIV>
__global__ void __launch_bounds__(kS3Threads, 1) stream3_kernel(const __grid_constant__ FusedArgs A, const TableLayout T,
                                                                unsigned char* __restrict__ ws) {
    constexpr int NT = kS3Threads, W = kS3Warps, PPT = kS3Ppt, TP = kS3TilePts, TF = kS3TileFloats;
    extern __shared__ __align__(128) unsigned char smem[];
    const msc_params& P = A.P;
    const FusedLayout& L = A.L;
    uint2* const cull = reinterpret_cast<uint2*>(smem + L.cull_off);            // .x box ids, .y edge code of the cull cell
    float* const boxp = reinterpret_cast<float*>(smem + L.boxp_off);            // [max_boxes][kBoxStride]
    uint32_t* const boxacc = reinterpret_cast<uint32_t*>(smem + L.boxacc_off);  // [max_boxes][kAccWords]
    S3Misc* const misc = reinterpret_cast<S3Misc*>(smem + L.misc_off);
    uint32_t* const wsinkc = reinterpret_cast<uint32_t*>(smem + L.window_off);  // [32 sink][W*W count][32 sink][W*W isum]
    uint32_t* const work_counter = reinterpret_cast<uint32_t*>(ws + T.counter_off);
    const float* const g_boxprep = reinterpret_cast<const float*>(ws + T.boxprep_off);
    const float* const g_wedges = reinterpret_cast<const float*>(ws + T.wedge_off);
    const uint32_t* const g_cullids = reinterpret_cast<const uint32_t*>(ws + T.cullids_off);

    const int tid = threadIdx.x, lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform as far as the compiler is concerned
    uint32_t smem_s = smem_u32(smem);
    asm volatile("" : "+r"(smem_s));  // opaque: one live register instead of a re-derived generic->shared conversion per use
    const uint32_t misc_s = smem_s + (uint32_t)L.misc_off;
    const uint32_t ring_s = smem_s + (uint32_t)L.tiles_off + (uint32_t)warp * (uint32_t)(2 * TF * 4);  // this warp's two slots
    const uint32_t bar_s = misc_s + (uint32_t)warp * 16u;                                               // its two mbarriers
    const uint32_t queue_s = smem_s + (uint32_t)L.queue_off + (uint32_t)warp * (64u * 16u);              // its candidate queue
    const uint32_t edge_s = misc_s + (uint32_t)offsetof(S3Misc, edge);
    const uint32_t cull_s = smem_s + (uint32_t)L.cull_off;
    const int win_w = L.win_w, win_lo = L.win_lo;
    const int n_win = win_w * win_w;
    const uint32_t sink_s = smem_s + (uint32_t)L.window_off + (uint32_t)lane * 4u;  // this lane's sink word of the count array
    const uint32_t wcount_s = smem_s + (uint32_t)L.window_off + 128u;
    const uint32_t isum_delta = (uint32_t)n_win * 4u + 128u;                        // count word -> isum word of the same cell (and sink -> sink)
    uint32_t* const wcount = wsinkc + 32;
    uint32_t* const wisum = wsinkc + 64 + n_win;
    const float* const ring = reinterpret_cast<const float*>(smem + L.tiles_off) + (size_t)warp * (2 * TF);
    const int res = P.bev_res, res_m1 = P.bev_res - 1;
    const size_t ncell = (size_t)res * (size_t)res;
    const int n_cams = P.n_cams;
    const int split = A.split;
    const uint64_t policy = l2_policy_evict_first();

    if (lane == 0) {
        mbar_init(&misc->full_bar[warp * 2 + 0], 1);
        mbar_init(&misc->full_bar[warp * 2 + 1], 1);
        mbar_fence_init();
    }
    uint32_t wk = 0;  // tiles this warp has consumed since launch: slot = wk & 1, mbarrier parity = (wk >> 1) & 1
    __syncthreads();

    for (;;) {
        // ------------------------------------------------------------ fetch a (sample, part) work item
        if (tid == 0) {
            const uint32_t w = atomicAdd(work_counter, 1u);
            const uint32_t s = w / (uint32_t)split;
            misc->sample = (int32_t)s;
            misc->part = (int32_t)(w - s * (uint32_t)split);
        }
        __syncthreads();
        const int sample = misc->sample, part = misc->part;
        if (sample >= A.in.n_samples) break;

        const int sw0 = A.in.sample_sweep_off[sample], sw1 = A.in.sample_sweep_off[sample + 1];
        const int n_sw = sw1 - sw0;
        if (tid < kS3PoseSmem && tid < n_sw) {
            misc->sweep_start[tid] = A.in.sweep_start[sw0 + tid];
            misc->sweep_count[tid] = A.in.sweep_count[sw0 + tid];
        }
        for (int i = tid; i < min(n_sw, kS3PoseSmem) * 12; i += NT) misc->pose[i] = A.in.sweep_pose[(size_t)sw0 * 12 + i];
        if (tid < MSC_STATS_STRIDE) misc->stats[tid] = 0u;
        if (FOV && tid < MSC_MAX_CAMS * 6) misc->wq[tid] = g_wedges[(size_t)sample * MSC_MAX_CAMS * 6 + tid];
        if (tid == 0) {
            misc->ci64 = reinterpret_cast<unsigned long long*>(A.out.bev_ci) + (size_t)sample * ncell;
            misc->h32 = reinterpret_cast<int*>(A.out.bev_height) + (size_t)sample * ncell;
        }
        __syncthreads();

        // Part `part` of the sample owns tiles (part * W + warp) + k * (split * W) of every sweep; warp `warp` streams them through its
        // private two-slot ring.  One cursor describes the tile in flight.
        const uint32_t tile_stride = (uint32_t)(split * W * TP), tile_first = (uint32_t)((part * W + warp) * TP);
        int n_si = -1;
        uint32_t n_first = 0, n_cnt = 0, n_base = 0;
        auto next_tile = [&]() -> bool {
            n_first += tile_stride;
            while (n_first >= n_cnt) {  // next sweep that still has a tile for this warp
                if (++n_si >= n_sw) return false;
                if (n_si < kS3PoseSmem) { n_cnt = misc->sweep_count[n_si]; n_base = misc->sweep_start[n_si]; }
                else { n_cnt = A.in.sweep_count[sw0 + n_si]; n_base = A.in.sweep_start[sw0 + n_si]; }
                n_first = tile_first;
            }
            return true;
        };
        auto issue = [&](uint32_t slot) {  // whole warp (uniform control flow); one lane talks to the TMA unit
            const uint32_t npts = min((uint32_t)TP, n_cnt - n_first);
            const uint32_t bytes = (npts * 20u + 15u) & ~15u;
            const float* src = A.in.points + ((size_t)n_base + n_first) * 5;
            if (lane == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s + slot * 8u), "r"(bytes) : "memory");
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                        ring_s + slot * (uint32_t)(TF * 4)),
                    "l"(src), "r"(bytes), "r"(bar_s + slot * 8u), "l"(policy)
                    : "memory");
            }
        };
        bool more = next_tile();
        if (more) issue(wk & 1u);  // overlaps the prologue below

        // ------------------------------------------------------------ prologue: accumulators, tables -> smem
        const int bx0 = A.in.sample_box_off[sample];
        int n_boxes = A.in.sample_box_off[sample + 1] - bx0;
        const bool box_overflow = n_boxes > L.max_boxes;  // caller under-declared max_boxes_per_sample
        if (box_overflow) n_boxes = L.max_boxes;
        const size_t cell_base = (size_t)sample * ncell;
        uint32_t* const g_ci = A.out.bev_ci + cell_base * 2;
        float* const g_h = A.out.bev_height + cell_base;
        const int n_cull = L.cull_dim * L.cull_dim;
        const int n_inner = L.inner_dim * L.inner_dim;
        uint32_t* const inner = reinterpret_cast<uint32_t*>(smem + L.inner_off);  // fine classes (one per BEV cell around the sensor)
        // class word of BEV cell (ix, iy): fine table inside [inner_lo, inner_lo + inner_dim)^2, else the cull cell's (low 24 bits of .y)
        auto class_of = [&](int ix, int iy) -> uint32_t {
            if (!FOV) return 0u;
            const uint32_t jx = (uint32_t)(ix - L.inner_lo), jy = (uint32_t)(iy - L.inner_lo);
            if (jx < (uint32_t)L.inner_dim && jy < (uint32_t)L.inner_dim) return inner[jy * (uint32_t)L.inner_dim + jx];
            return cull[(iy >> L.cull_shift) * L.cull_dim + (ix >> L.cull_shift)].y & 0xffffffu;
        };
        {
            const uint32_t* ids = g_cullids + (size_t)sample * n_cull;       // candidate boxes per cull cell (fused_cullids_kernel)
            // edge classes of the sample's camera wedges per cull cell, and per BEV cell around the sensor (fine table)
            for (int i = tid; i < n_cull; i += NT) {
                const uint32_t cls = FOV ? edge_class_word(A, misc->wq, i) : 0u;
                cull[i] = make_uint2(ids[i], cls | (s3_code_of(cls) << kCodeShift));
            }
            if (FOV)
                for (int i = tid; i < n_inner; i += NT) inner[i] = edge_class_word(A, misc->wq, n_cull + i);
            if (tid < 32) { wsinkc[tid] = 0u; wsinkc[32 + n_win + tid] = 0u; }
            for (int i = tid; i < n_win; i += NT) wisum[i] = 0u;
            for (int i = tid; i < n_boxes * kAccWords; i += NT) boxacc[i] = ((i % kAccWords) == 1) ? 0x7f800000u : 0u;
            const float4* bsrc = reinterpret_cast<const float4*>(g_boxprep + (size_t)bx0 * kBoxStride);
            for (int i = tid; i < n_boxes * (kBoxStride / 4); i += NT) reinterpret_cast<float4*>(boxp)[i] = bsrc[i];
            if (tid < 32) {
                S3Edge E;
                E.ax = 0.0f; E.ay = 0.0f; E.a = __int_as_float(0x7fc00000); E.b = 0.0f;  // pad: the test fails
                E.inc_lo = E.inc_hi = E.pad0 = E.pad1 = 0u;
                const int e = tid - 1, c = e & (MSC_MAX_CAMS - 1);
                if (FOV && e >= 0 && e < 2 * MSC_MAX_CAMS && c < n_cams) {
                    const float* wq = g_wedges + ((size_t)sample * MSC_MAX_CAMS + c) * 6;
                    // right edge: fma(w4, qy, -(w5 * qx)) >= 0; left edge: fma(w3, qx, -(w2 * qy)) >= 0  (in_wedge)
                    E.ax = wq[0]; E.ay = wq[1];
                    if (e >= MSC_MAX_CAMS) { E.a = wq[3]; E.b = wq[2]; } else { E.a = wq[4]; E.b = wq[5]; }
                    E.inc_lo = c < 4 ? 1u << (8 * c) : 0u;
                    E.inc_hi = c < 4 ? 0u : 1u << (8 * (c - 4));
                }
                misc->edge[tid] = E;
            }
            if (split == 1) {  // zero-fill this sample's global layers (a split sample is zero-filled by the host before the launch)
                uint4* c4 = reinterpret_cast<uint4*>(g_ci);
                for (size_t i = tid; i < ncell / 2; i += NT) c4[i] = make_uint4(0, 0, 0, 0);
                uint4* h4 = reinterpret_cast<uint4*>(g_h);
                for (size_t i = tid; i < ncell / 4; i += NT) h4[i] = make_uint4(0, 0, 0, 0);
            }
        }
        __threadfence();
        __syncthreads();  // class tables are in smem
        for (int i = tid; i < n_win; i += NT) {
            const int wy = i / win_w, wx = i - wy * win_w;
            wcount[i] = s3_code_of(class_of(wx + win_lo, wy + win_lo)) << kCodeShift;
        }
        __syncthreads();

        // ------------------------------------------------------------ main loop: this warp's tiles, no cross-warp sync
        uint32_t c_removed = 0, c_ground = 0, c_periph = 0;  // per-thread counters (flushed once per sample)
        uint32_t cam_lo = 0, cam_hi = 0;       // eight 8-bit per-camera counters of the exact edge tests, spilled every <= 255 points
        uint32_t cam_pts = 0;
        uint32_t q_head = 0, q_tail = 0;  // warp-uniform (every lane derives them from the same ballots); < 32 pending between points
        // every lane tests one queued point against its candidate boxes; the (usually single) containing box is accumulated
        // once after the loop, a second containing box (overlapping annotations) inside it
        auto drain_queue = [&](uint32_t n_take) {
            const float4 e = s3_lds128(queue_s + (((q_head + (uint32_t)lane) & 63u) << 4));
            if ((uint32_t)lane < n_take) {
                uint32_t ids = __float_as_uint(e.w);
                const float es2 = __fadd_rn(__fmul_rn(e.x, e.x), __fmul_rn(e.y, e.y));
                const uint32_t fx = (uint32_t)(__float2int_rn(__fmul_rn(e.x, A.cscale)) + A.centroid_bias);
                const uint32_t fy = (uint32_t)(__float2int_rn(__fmul_rn(e.y, A.cscale)) + A.centroid_bias);
                const uint32_t fz = (uint32_t)(__float2int_rn(__fmul_rn(e.z, A.cscale)) + A.centroid_bias);
                // centroid sums: one 32-bit word per axis plus a carry word that takes a rare second atomic when the word wraps
                // (the coordinate is a 24-bit value, so that is at most once per 256 points); still order-independent integers
                auto accumulate = [&](int b) {
                    const uint32_t acc_s = smem_s + (uint32_t)L.boxacc_off + (uint32_t)b * (uint32_t)(kAccWords * 4);
                    s3_red_add(acc_s, 1u);
                    asm volatile("red.shared.min.u32 [%0], %1;" ::"r"(acc_s + 4u), "r"(__float_as_uint(es2)) : "memory");
                    const uint32_t ox = s3_atom_add(acc_s + 8u, fx), oy = s3_atom_add(acc_s + 12u, fy), oz = s3_atom_add(acc_s + 16u, fz);
                    if (ox > ~fx) s3_red_add(acc_s + 20u, 1u);
                    if (oy > ~fy) s3_red_add(acc_s + 24u, 1u);
                    if (oz > ~fz) s3_red_add(acc_s + 28u, 1u);
                };
                int hit = -1;
                if (ids == kCullAll) {  // crowded cell (more than four boxes): test every box
                    for (int b = 0; b < n_boxes; ++b)
                        if (box_contains(boxp, b, e.x, e.y, e.z)) {
                            if (hit >= 0) accumulate(b); else hit = b;
                        }
                } else {
                    do {
                        const int b = (int)(ids & 0xffu);
                        if (box_contains(boxp, b, e.x, e.y, e.z)) {
                            if (hit >= 0) accumulate(b); else hit = b;
                        }
                        ids = (ids >> 8) | 0xff000000u;
                    } while ((ids & 0xffu) != 0xffu);
                }
                if (hit >= 0) accumulate(hit);
            }
            q_head += n_take;
        };
        uint32_t pose_s = 0;
        int pose_si = -1;
        while (more) {
            // ---- the tile in flight becomes the current one; its successor takes the slot consumed in the previous iteration
            const uint32_t npts = min((uint32_t)TP, n_cnt - n_first);
            if (n_si != pose_si) {
                pose_si = n_si;
                if (n_si < kS3PoseSmem) {
                    pose_s = misc_s + (uint32_t)offsetof(S3Misc, pose) + (uint32_t)n_si * 96u;
                } else {
                    __syncwarp();
                    if (lane < 12) misc->wpose[warp * 12 + lane] = A.in.sweep_pose[(size_t)(sw0 + n_si) * 12 + lane];
                    __syncwarp();
                    pose_s = misc_s + (uint32_t)offsetof(S3Misc, wpose) + (uint32_t)warp * 96u;
                }
            }
            more = next_tile();
            if (more) issue((wk + 1u) & 1u);
            {
                const uint32_t bar = bar_s + (wk & 1u) * 8u, parity = (wk >> 1) & 1u;
                asm volatile(
                    "{\n"
                    ".reg .pred p;\n"
                    "MSC_S3WAIT_%=:\n"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                    "@p bra MSC_S3DONE_%=;\n"
                    "bra MSC_S3WAIT_%=;\n"
                    "MSC_S3DONE_%=:\n"
                    "}\n" ::"r"(bar),
                    "r"(parity)
                    : "memory");
            }
            const float* tp = ring + (wk & 1u) * TF + lane * 5;

            // ---- phase A: branch-free over the lane's points so their dependency chains interleave
            float xr[PPT], yr[PPT], zr[PPT], inten[PPT];
            bool keep[PPT];
            {
                float x[PPT], y[PPT], z[PPT];
#pragma unroll
                for (int u = 0; u < PPT; ++u) {
                    x[u] = tp[u * 160 + 0]; y[u] = tp[u * 160 + 1]; z[u] = tp[u * 160 + 2];
                    inten[u] = tp[u * 160 + 3];
                }
                if (npts < (uint32_t)TP) {  // last tile of a sweep: rows past its end hold stale data -> NaN fails every compare below
#pragma unroll
                    for (int u = 0; u < PPT; ++u)
                        if ((uint32_t)lane + (uint32_t)u * 32u >= npts) x[u] = __int_as_float(0x7fc00000);
                }
                double xd[PPT], yd[PPT], zd[PPT];
#pragma unroll
                for (int u = 0; u < PPT; ++u) {
                    // A.1 remove_close (square, sweep's own sensor frame)
                    const bool close = (fabsf(x[u]) < P.remove_close_radius) & (fabsf(y[u]) < P.remove_close_radius);
                    keep[u] = !close;
                    if (close) ++c_removed;
                    xd[u] = (double)x[u]; yd[u] = (double)y[u]; zd[u] = (double)z[u];
                }
                __syncwarp();  // every lane has read its rows: the slot may be refilled at the top of the next iteration
                // A.1 f64 matrix x f32 point -> f32, one matrix row at a time
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    double m0, m1, m2, m3;
                    s3_lds_f64x2(pose_s + r * 32, m0, m1);
                    s3_lds_f64x2(pose_s + r * 32 + 16, m2, m3);
#pragma unroll
                    for (int u = 0; u < PPT; ++u) {
                        const float v = (float)__fma_rn(m0, xd[u], __fma_rn(m1, yd[u], __fma_rn(m2, zd[u], m3)));
                        if (r == 0) xr[u] = v; else if (r == 1) yr[u] = v; else zr[u] = v;
                    }
                }
            }
            // ---- phase B: filter, BEV cell, cull entry, window word
            uint32_t cand[PPT], code[PPT], cell[PPT], q[PPT];
            bool periph[PPT];
            bool rare = false, wide = false;
#pragma unroll
            for (int u = 0; u < PPT; ++u) {
                // lidar_agent.py:106-110, sqrt-free (thresholds on s are exact, geometry.sqrt_thresholds)
                const float s2 = __fadd_rn(__fmul_rn(xr[u], xr[u]), __fmul_rn(yr[u], yr[u]));
                keep[u] = keep[u] && (s2 >= P.s_lo) && (s2 <= P.s_hi) && (zr[u] < P.z_max) && (zr[u] > P.z_min);
                // BEV cell, lidar_agent.py:547-552 (garbage for dropped points is clamped and never used)
                int ix, iy;
                bev_cell_xy<FASTDIV>(xr[u], yr[u], P.bev_range, A.two_r, A.rcp_two_r, A.resf, res_m1, ix, iy);
                const uint2 ce = s3_lds64(cull_s + (uint32_t)(((iy >> L.cull_shift) * L.cull_dim + (ix >> L.cull_shift)) << 3));
                const uint32_t wx = (uint32_t)(ix - win_lo), wy = (uint32_t)(iy - win_lo);
                const bool inwin = max(wx, wy) < (uint32_t)win_w;
                // ---- phase C: the window atomic counts the point and returns the cell's edge code; dropped and out-of-window
                // points count into the lane's sink word (code 0) instead of branching
                const uint32_t wa = (keep[u] && inwin) ? wcount_s + ((wy * (uint32_t)win_w + wx) << 2) : sink_s;
                const uint32_t old = s3_atom_add(wa, 1u);
                // Q8 intensity, clamp [0, 65535], NaN -> 0: fmaxf drops NaN and negatives, the FFMA rounds v * 2^shift to nearest even in
                // the low mantissa bits of 2^23 + v * 2^shift (exact while below 2^23; larger values clamp anyway)
                q[u] = min(__float_as_uint(__fmaf_rn(fmaxf(inten[u], 0.0f), A.iscale, 8388608.0f)) - 0x4b000000u, 65535u);
                s3_red_add(wa + isum_delta, q[u]);
                if (keep[u] && zr[u] < P.ground_z) ++c_ground;  // lidar_agent.py:128
                periph[u] = keep[u] && !inwin;
                code[u] = (periph[u] ? ce.y : old) >> kCodeShift;
                cand[u] = keep[u] ? ce.x : kCullEmpty;
                cell[u] = ((uint32_t)iy << 16) | (uint32_t)ix;  // (the linear index is formed by the few lanes that need it)
                wide = wide || periph[u] || (keep[u] && zr[u] > 0.0f);
                rare = rare || (code[u] == kCodeMulti);
            }
            if (wide) {  // a few lanes per warp: cells outside the window (one 64-bit RED) and the max-height layer (z > 0 only)
                unsigned long long* const ci64 = *reinterpret_cast<unsigned long long* volatile*>(&misc->ci64);
                int* const h32 = *reinterpret_cast<int* volatile*>(&misc->h32);
#pragma unroll
                for (int u = 0; u < PPT; ++u) {
                    const uint32_t lin = (cell[u] >> 16) * (uint32_t)res + (cell[u] & 0xffffu);
                    if (periph[u]) {
                        asm volatile("red.global.add.u64 [%0], %1;" ::"l"(ci64 + lin), "l"(1ull | ((unsigned long long)q[u] << 32)) : "memory");
                        ++c_periph;
                        if (FOV) {  // cameras whose wedge holds the whole cull cell (window cells get theirs at the flush)
                            const uint32_t d = s3_decided_in(cull[((cell[u] >> 16) >> L.cull_shift) * L.cull_dim + ((cell[u] & 0xffffu) >> L.cull_shift)].y);
                            cam_lo += ((d & 0xfu) * 0x00204081u) & 0x01010101u;
                            cam_hi += ((d >> 4) * 0x00204081u) & 0x01010101u;
                        }
                    }
                    if (keep[u] && zr[u] > 0.0f)  // :560, 0-initialised max
                        asm volatile("red.global.max.s32 [%0], %1;" ::"l"(h32 + lin), "r"(__float_as_int(zr[u])) : "memory");
                }
            }
            // ---- phase F: one exact cross product for the points of cells a single image-column ray crosses (branch-free: every other
            // point reads a pad entry whose test fails); the camera's byte counter takes the entry's increment
            if (FOV) {
#pragma unroll
                for (int u = 0; u < PPT; ++u) {
                    const uint32_t es = edge_s + (code[u] << 5);
                    const float4 E = s3_lds128(es);
                    const uint2 inc = s3_lds64(es + 16u);
                    const float qx = __fsub_rn(xr[u], E.x), qy = __fsub_rn(yr[u], E.y);
                    const bool left = code[u] > (uint32_t)MSC_MAX_CAMS;
                    const float sv = left ? qx : qy, tv = left ? qy : qx;
                    const float cr = __fmaf_rn(E.z, sv, -__fmul_rn(E.w, tv));
                    if (cr >= 0.0f) { cam_lo += inc.x; cam_hi += inc.y; }
                }
                if (__any_sync(0xffffffffu, rare)) {  // cold: cells crossed by two or more rays (next to the cameras)
                    bool m0 = code[0] == kCodeMulti, m1 = code[1] == kCodeMulti;
                    static_assert(PPT == 2, "the cold path selects between two points");
                    while (m0 || m1) {  // one of the lane's multi-edge points per trip (a lane rarely has two)
                        const float px = m0 ? xr[0] : xr[1], py = m0 ? yr[0] : yr[1];
                        const uint32_t cxy = m0 ? cell[0] : cell[1];
                        const bool pp = m0 ? periph[0] : periph[1];
                        if (m0) m0 = false; else m1 = false;
                        // the class the point's code came from: its cull cell's outside the window, else its BEV cell's
                        const int cix = (int)(cxy & 0xffffu), ciy = (int)(cxy >> 16);
                        const uint32_t cls = pp ? (cull[(ciy >> L.cull_shift) * L.cull_dim + (cix >> L.cull_shift)].y & 0xffffffu) : class_of(cix, ciy);
                        uint32_t und = (cls >> 8) & 0xffffu, pass = 0xffffu;
                        while (und) {
                            const int e = __ffs((int)und) - 1;
                            und &= und - 1u;
                            const float4 E = s3_lds128(edge_s + ((uint32_t)(e + 1) << 5));
                            const float qx = __fsub_rn(px, E.x), qy = __fsub_rn(py, E.y);
                            const bool left = e >= MSC_MAX_CAMS;
                            const float sv = left ? qx : qy, tv = left ? qy : qx;
                            const float cr = __fmaf_rn(E.z, sv, -__fmul_rn(E.w, tv));
                            if (!(cr >= 0.0f)) pass &= ~(1u << e);
                        }
                        // cameras with an undecided edge in this cell whose every undecided edge passed
                        const uint32_t any_und = ((cls >> 8) | (cls >> 16)) & 0xffu;
                        const uint32_t in = cls & any_und & pass & (pass >> 8);
                        cam_lo += ((in & 0xfu) * 0x00204081u) & 0x01010101u;
                        cam_hi += ((in >> 4) * 0x00204081u) & 0x01010101u;
                    }
                }
            }
            // ---- phase D: points that have candidate boxes go to this warp's queue; whenever 32 are pending every lane tests
            // one of them (dense), instead of a handful of lanes looping while the rest of the warp idles
#pragma unroll
            for (int u = 0; u < PPT; ++u) {
                const bool has = cand[u] != kCullEmpty;
                const uint32_t m = __ballot_sync(0xffffffffu, has);
                if (has) {
                    uint32_t lt_mask;
                    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
                    const uint32_t slot = (q_tail + __popc(m & lt_mask)) & 63u;
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(queue_s + (slot << 4)), "f"(xr[u]), "f"(yr[u]), "f"(zr[u]),
                                 "f"(__uint_as_float(cand[u]))
                                 : "memory");
                }
                q_tail += __popc(m);
                if (q_tail - q_head >= 32u) {
                    __syncwarp();  // the entries stored above are visible to the lanes that test them
                    drain_queue(32u);
                }
            }
            ++wk;
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
        while (q_tail != q_head) {
            __syncwarp();
            drain_queue(min(q_tail - q_head, 32u));
        }

        // ------------------------------------------------------------ epilogue
        __syncthreads();  // every tile of this part is accumulated
        // Everything that leaves the CTA is additive (or a min): per-thread counters, then the window -- its cells go to the global layers,
        // their counts give n_kept and, times the cameras that hold a whole cell, the decided share of the per-camera counts.
        uint32_t kept = c_periph, flags = 0, cam[MSC_MAX_CAMS];
#pragma unroll
        for (int c = 0; c < MSC_MAX_CAMS; ++c) cam[c] = FOV ? (((c < 4 ? cam_lo : cam_hi) >> ((c & 3) * 8)) & 0xffu) : 0u;
        unsigned long long* const ci64 = reinterpret_cast<unsigned long long*>(g_ci);
        {
            const int half_w = win_w >> 1;  // win_w and win_lo are even -> 16-byte aligned rows of the global layer
            for (int i = tid; i < win_w * half_w; i += NT) {
                const int wy = i / half_w, wx = (i - wy * half_w) * 2;
                const uint2 c2 = *reinterpret_cast<const uint2*>(wcount + wy * win_w + wx);
                const uint2 s2 = *reinterpret_cast<const uint2*>(wisum + wy * win_w + wx);
                const uint32_t c0 = c2.x & kCountMask, c1 = c2.y & kCountMask;
                const int ix = wx + win_lo, iy = wy + win_lo;
                const size_t cell = (size_t)iy * (size_t)res + (size_t)ix;
                if (split == 1) {
                    *reinterpret_cast<uint4*>(g_ci + cell * 2) = make_uint4(c0, s2.x, c1, s2.y);
                } else {  // a part of a split sample merges with reductions (the host zero-filled the layers)
                    if (c0) atomicAdd(ci64 + cell, (unsigned long long)c0 | ((unsigned long long)s2.x << 32));
                    if (c1) atomicAdd(ci64 + cell + 1, (unsigned long long)c1 | ((unsigned long long)s2.y << 32));
                }
                if ((c0 | c1) == 0u) continue;
                kept += c0 + c1;
                flags |= (c0 >= 65536u || c1 >= 65536u) ? 1u : 0u;
                if (FOV) {
                    const uint32_t d0 = s3_decided_in(class_of(ix, iy)), d1 = s3_decided_in(class_of(ix + 1, iy));
#pragma unroll
                    for (int c = 0; c < MSC_MAX_CAMS; ++c) cam[c] += (((d0 >> c) & 1u) ? c0 : 0u) + (((d1 >> c) & 1u) ? c1 : 0u);
                }
            }
        }
        {
            uint32_t v[4 + MSC_MAX_CAMS];
            v[0] = c_removed; v[1] = kept; v[2] = c_ground; v[3] = 0u;
#pragma unroll
            for (int c = 0; c < MSC_MAX_CAMS; ++c) v[4 + c] = cam[c];
#pragma unroll
            for (int i = 0; i < 4 + (FOV ? MSC_MAX_CAMS : 0); ++i) {
                if (i == 3) continue;
                const uint32_t r = __reduce_add_sync(0xffffffffu, v[i]);
                if (lane == 0 && r) atomicAdd(&misc->stats[1 + i], r);  // [1] removed, [2] kept, [3] ground, [5 + c] per camera
            }
            const uint32_t f = __reduce_or_sync(0xffffffffu, flags);
            if (lane == 0 && f) atomicOr(&misc->stats[13], f);
        }
        __syncthreads();
        uint32_t* const g_stats = A.out.stats + (size_t)sample * MSC_STATS_STRIDE;
        bool finalise = true;
        if (split > 1) {
            unsigned long long* const scr = reinterpret_cast<unsigned long long*>(ws + T.boxscr_off) + (size_t)bx0 * 4;
            for (int b = tid; b < n_boxes; b += NT) {
                const uint32_t* acc = boxacc + b * kAccWords;
                if (acc[0]) {
                    uint32_t* s32 = reinterpret_cast<uint32_t*>(scr + (size_t)b * 4);
                    atomicAdd(s32, acc[0]);
                    atomicMin(s32 + 1, acc[1]);
#pragma unroll
                    for (int k = 0; k < 3; ++k) atomicAdd(scr + (size_t)b * 4 + 1 + k, (unsigned long long)acc[2 + k] + ((unsigned long long)acc[5 + k] << 32));
                }
            }
            uint32_t* const gsc = reinterpret_cast<uint32_t*>(ws + T.splitstats_off) + (size_t)sample * MSC_STATS_STRIDE;
            if (tid < MSC_STATS_STRIDE && tid != 15 && misc->stats[tid]) {
                if (tid == 13) atomicOr(gsc + tid, misc->stats[tid]); else atomicAdd(gsc + tid, misc->stats[tid]);
            }
            // the part that takes the last ticket sees every other part's reductions and finalises the sample
            __threadfence();
            __syncthreads();
            if (tid == 0) misc->ticket = (int32_t)atomicAdd(gsc + 15, 1u);
            __syncthreads();
            finalise = misc->ticket == split - 1;
            if (finalise) {
                __threadfence();
                if (tid < MSC_STATS_STRIDE && tid != 15) misc->stats[tid] = __ldcg(gsc + tid);
            }
        }
        if (finalise) {
            const unsigned long long* const scr = reinterpret_cast<const unsigned long long*>(ws + T.boxscr_off) + (size_t)bx0 * 4;
            for (int b = tid; b < n_boxes; b += NT) {
                uint32_t cnt, mn;
                unsigned long long sum3[3];
                if (split == 1) {
                    const uint32_t* acc = boxacc + b * kAccWords;
                    cnt = acc[0]; mn = acc[1];
#pragma unroll
                    for (int k = 0; k < 3; ++k) sum3[k] = (unsigned long long)acc[2 + k] + ((unsigned long long)acc[5 + k] << 32);
                } else {
                    const unsigned long long w0 = __ldcg(scr + (size_t)b * 4);
                    cnt = (uint32_t)w0; mn = (uint32_t)(w0 >> 32);
#pragma unroll
                    for (int k = 0; k < 3; ++k) sum3[k] = __ldcg(scr + (size_t)b * 4 + 1 + k);
                }
                const size_t o = (size_t)(bx0 + b);
                A.out.box_count[o] = cnt;
                if (cnt == 0) {
                    A.out.box_nearest[o] = INFINITY;
                    A.out.box_centroid[o * 3 + 0] = 0.0f; A.out.box_centroid[o * 3 + 1] = 0.0f; A.out.box_centroid[o * 3 + 2] = 0.0f;
                } else {
                    A.out.box_nearest[o] = __fsqrt_rn(__uint_as_float(mn));
                    const double den = (double)cnt * (double)A.cscale;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const long long sum = (long long)sum3[k] - (long long)cnt * (long long)A.centroid_bias;
                        A.out.box_centroid[o * 3 + k] = (float)((double)sum / den);
                    }
                }
            }
            __syncthreads();  // (split > 1: misc->stats holds the merged counters)
            if (tid < MSC_STATS_STRIDE) {
                uint32_t v = misc->stats[tid];
                uint32_t n_in = 0;
                if (tid <= 1)
                    for (int s = sw0; s < sw1; ++s) n_in += A.in.sweep_count[s];
                if (tid == 0) v = n_in;
                if (tid == 1) v = n_in - v;                            // n_after_close = n_in - removed
                if (tid == 4) v = misc->stats[2] - misc->stats[3];     // n_object = n_kept - n_ground
                if (tid == 13 && box_overflow) v |= 0x80000000u;
                if (tid == 15) v = 0u;
                g_stats[tid] = v;
            }
        }
        // (the __syncthreads after the next work-item fetch orders these reads before the smem is re-initialised)
    }
}

template <bool FOV, bool FASTDIV>
static int s3_launch_one(const FusedArgs& args, const TableLayout& T, unsigned char* ws, int grid, cudaStream_t stream) {
    auto kern = stream3_kernel<FOV, FASTDIV>;
    MSC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, args.L.total_bytes));
    kern<<<grid, kS3Threads, args.L.total_bytes, stream>>>(args, T, ws);
    MSC_CUDA(cudaGetLastError());
    return MSC_OK;
}

int launch_stream3_kernel(const FusedArgs& args, const TableLayout& T, unsigned char* ws, int grid, bool fov, bool fast, cudaStream_t stream) {
    if (fov) return fast ? s3_launch_one<true, true>(args, T, ws, grid, stream) : s3_launch_one<true, false>(args, T, ws, grid, stream);
    return fast ? s3_launch_one<false, true>(args, T, ws, grid, stream) : s3_launch_one<false, false>(args, T, ws, grid, stream);
}

}  // namespace msc
