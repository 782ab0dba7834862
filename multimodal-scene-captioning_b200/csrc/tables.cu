// tables.cu -- small-table kernels: per-annotation table (scenegraph_agent.py:186-225, :281-295;
// baseline_gpt4o.py:304-317), box footprints, the [EXT] pairwise relation table and the standalone
// box -> camera projection (App. A.3).  All float64 with IEEE + - * / sqrt and no FMA (-fmad=false), so the
// numeric columns are bit-identical to the scalar definition; category columns use exact predicates.
#include "msc_common.cuh"

namespace msc {

// zones in the order of scenegraph_agent.py:136-146
__constant__ double c_zone_min[9] = {0, 10, 30, 0, 10, 0, 10, 0, 10};
__constant__ double c_zone_max[9] = {10, 30, 50, 10, 30, 10, 30, 10, 30};
__constant__ uint8_t c_zone_dir[9] = {0, 0, 0, 1, 1, 3, 3, 2, 2};

__global__ void __launch_bounds__(256) annotation_table_kernel(int n, const double* __restrict__ xy, const double* __restrict__ vel,
                                                              double* __restrict__ distance, uint8_t* __restrict__ direction,
                                                              uint8_t* __restrict__ moving, uint8_t* __restrict__ zone,
                                                              uint8_t* __restrict__ region_bits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = xy[2 * i], y = xy[2 * i + 1];
    const double d = sqrt(x * x + y * y);  // :189
    distance[i] = d;
    const uint8_t dir = bearing4(x, y);    // :190-201
    direction[i] = dir;
    const double vx = vel[2 * i], vy = vel[2 * i + 1];
    const double sp = sqrt(vx * vx + vy * vy);  // :216
    moving[i] = (sp > 0.5) ? 1 : 0;             // :217 (NaN compares false -> stopped)
    uint8_t zc = 255;
#pragma unroll
    for (int k = 8; k >= 0; --k)
        if (dir == c_zone_dir[k] && c_zone_min[k] <= d && d < c_zone_max[k]) zc = (uint8_t)k;  // first match wins (:290-293)
    zone[i] = zc;
    region_bits[i] = (uint8_t)((x > 0.0 ? 1 : 0) | (y > 0.0 ? 2 : 0));  // baseline_gpt4o.py:309-317
}

__global__ void __launch_bounds__(128) box_footprints_kernel(int n, const double* __restrict__ boxes, const double* __restrict__ ego_pose,
                                                            double* __restrict__ rect) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const double* box = boxes + (size_t)b * 10;
    double c[3] = {box[0], box[1], box[2]};
    double R[9];
    quat_to_rot(box + 6, R);
    if (ego_pose) frame_change(ego_pose, c, R);
    double ux = R[0], uy = R[3];
    const double nn = sqrt(ux * ux + uy * uy);
    if (nn > 0.0) { ux = ux / nn; uy = uy / nn; } else { ux = 1.0; uy = 0.0; }
    double* o = rect + (size_t)b * 6;
    o[0] = c[0]; o[1] = c[1]; o[2] = ux; o[3] = uy; o[4] = box[4] / 2.0; o[5] = box[3] / 2.0;
}

// one thread per ordered pair; rect rows are staged through smem per 16x16 tile
// blockIdx.z selects the sample of a batch: sample s uses rect rows [off[s], off[s+1]) and writes its n_s x n_s tables at pair_off[s]
__global__ void __launch_bounds__(256) relation_table_kernel(int n_single, const int32_t* __restrict__ box_off, const int64_t* __restrict__ pair_off,
                                                            const double* __restrict__ rect, float* __restrict__ dist,
                                                            float* __restrict__ bearing, uint8_t* __restrict__ category,
                                                            uint8_t* __restrict__ overlap) {
    __shared__ double sA[16][6], sB[16][6];
    int n = n_single;
    if (box_off) {
        const int s = blockIdx.z;
        n = box_off[s + 1] - box_off[s];
        rect += (size_t)box_off[s] * 6;
        const size_t po = (size_t)pair_off[s];
        dist += po; bearing += po; category += po; overlap += po;
        if ((int)(blockIdx.x * 16) >= n || (int)(blockIdx.y * 16) >= n) return;  // whole tile outside this sample (uniform per block)
    }
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int j = blockIdx.x * 16 + tx, i = blockIdx.y * 16 + ty;
    if (threadIdx.x < 96) {
        const int r = threadIdx.x / 6, k = threadIdx.x - r * 6;
        const int gi = blockIdx.y * 16 + r, gj = blockIdx.x * 16 + r;
        sA[r][k] = (gi < n) ? rect[(size_t)gi * 6 + k] : 0.0;
        sB[r][k] = (gj < n) ? rect[(size_t)gj * 6 + k] : 0.0;
    }
    __syncthreads();
    if (i >= n || j >= n) return;
    const double* A = sA[ty];
    const double* B = sB[tx];
    const size_t o = (size_t)i * n + j;
    const double dx = B[0] - A[0], dy = B[1] - A[1];
    dist[o] = (float)sqrt(dx * dx + dy * dy);
    double ang = atan2(dy, dx) * 180.0 / 3.14159265358979323846;  // scenegraph_agent.py:190
    ang = fmod(ang + 360.0, 360.0);                               // :191
    bearing[o] = (float)ang;
    category[o] = bearing4(dx, dy);
    bool sep = false;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double ex = (k == 0) ? A[2] : (k == 1) ? -A[3] : (k == 2) ? B[2] : -B[3];
        const double ey = (k == 0) ? A[3] : (k == 1) ? A[2] : (k == 2) ? B[3] : B[2];
        const double t = fabs(dx * ex + dy * ey);
        const double ra = A[4] * fabs(A[2] * ex + A[3] * ey) + A[5] * fabs(-A[3] * ex + A[2] * ey);
        const double rb = B[4] * fabs(B[2] * ex + B[3] * ey) + B[5] * fabs(-B[3] * ex + B[2] * ey);
        if (t > ra + rb) sep = true;
    }
    overlap[o] = sep ? 0 : 1;
}

__global__ void __launch_bounds__(128) project_boxes_kernel(int n_boxes, const double* __restrict__ boxes, int n_cams,
                                                           const double* __restrict__ cam_ego_pose, const double* __restrict__ cam_calib,
                                                           const double* __restrict__ cam_K, double W, double H, uint8_t* __restrict__ visible,
                                                           float* __restrict__ extent) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_boxes * n_cams) return;
    const int b = t / n_cams, c = t - b * n_cams;
    project_box(boxes + (size_t)b * 10, cam_ego_pose + (size_t)c * 7, cam_calib + (size_t)c * 7, cam_K + (size_t)c * 9, W, H, visible + t,
                extent + (size_t)t * 4);
}

}  // namespace msc

extern "C" {

int msc_annotation_table(int32_t n, const double* xy, const double* vel, double* distance, uint8_t* direction, uint8_t* moving,
                         uint8_t* zone, uint8_t* region_bits, void* stream) {
    using namespace msc;
    MSC_REQUIRE(n >= 0, "negative n");
    if (n == 0) return MSC_OK;
    MSC_REQUIRE(xy && vel && distance && direction && moving && zone && region_bits, "null argument");
    annotation_table_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, xy, vel, distance, direction, moving, zone, region_bits);
    MSC_CUDA(cudaGetLastError());
    return MSC_OK;
}

int msc_box_footprints(int32_t n, const double* boxes, const double* ego_pose, double* rect, void* stream) {
    using namespace msc;
    MSC_REQUIRE(n >= 0, "negative n");
    if (n == 0) return MSC_OK;
    MSC_REQUIRE(boxes && rect, "null argument");
    box_footprints_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n, boxes, ego_pose, rect);
    MSC_CUDA(cudaGetLastError());
    return MSC_OK;
}

int msc_relation_table(int32_t n, const double* rect, float* dist, float* bearing, uint8_t* category, uint8_t* overlap, void* stream) {
    using namespace msc;
    MSC_REQUIRE(n >= 0, "negative n");
    if (n == 0) return MSC_OK;
    MSC_REQUIRE(rect && dist && bearing && category && overlap, "null argument");
    dim3 grid((n + 15) / 16, (n + 15) / 16);
    relation_table_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, nullptr, nullptr, rect, dist, bearing, category, overlap);
    MSC_CUDA(cudaGetLastError());
    return MSC_OK;
}

int msc_relation_table_batch(int32_t n_samples, int32_t max_boxes, const int32_t* box_off, const int64_t* pair_off, const double* rect, float* dist,
                             float* bearing, uint8_t* category, uint8_t* overlap, void* stream) {
    using namespace msc;
    MSC_REQUIRE(n_samples >= 0 && max_boxes >= 0, "negative counts");
    if (n_samples == 0 || max_boxes == 0) return MSC_OK;
    MSC_REQUIRE(n_samples <= 65535, "at most 65535 samples per call");
    MSC_REQUIRE(box_off && pair_off && rect && dist && bearing && category && overlap, "null argument");
    dim3 grid((max_boxes + 15) / 16, (max_boxes + 15) / 16, n_samples);
    relation_table_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(0, box_off, pair_off, rect, dist, bearing, category, overlap);
    MSC_CUDA(cudaGetLastError());
    return MSC_OK;
}

int msc_project_boxes(int32_t n_boxes, const double* boxes, int32_t n_cams, const double* cam_ego_pose, const double* cam_calib,
                      const double* cam_K, int32_t image_w, int32_t image_h, uint8_t* visible, float* extent, void* stream) {
    using namespace msc;
    MSC_REQUIRE(n_boxes >= 0 && n_cams >= 0 && n_cams <= MSC_MAX_CAMS, "bad counts");
    if (n_boxes == 0 || n_cams == 0) return MSC_OK;
    MSC_REQUIRE(boxes && cam_ego_pose && cam_calib && cam_K && visible && extent, "null argument");
    const int t = n_boxes * n_cams;
    project_boxes_kernel<<<(t + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n_boxes, boxes, n_cams, cam_ego_pose, cam_calib, cam_K,
                                                                           (double)image_w, (double)image_h, visible, extent);
    MSC_CUDA(cudaGetLastError());
    return MSC_OK;
}

}  // extern "C"
