// tdf.cu — truncated (squared) distance field of a keypoint's occupied-voxel list: the reference's only CUDA
// (RealTimeRobot/kernel.cu:8-31 ComputeTDF, :34-106 ComputeTDFWithCuda; called from KeyPoint::get_TSDF, key_point.h:313).
//
//   out[z*dim*dim + y*dim + x] = min(900, min_i (x-xi)^2 + (y-yi)^2 + (z-zi)^2)      integer voxel units, stored fp32
//
// B200 form: all keypoints of a cloud in ONE launch (grid.y = keypoint), the occupied list staged through shared
// memory in tiles (every thread of a CTA reads every triple: smem broadcast instead of num_occ global loads per voxel),
// integer min kept in a register, no per-call cudaMalloc / cudaFree / cudaDeviceSynchronize.  Bit-exact with the
// reference: the arithmetic is int32 and the result (<= 900) is exactly representable.
#include "common.cuh"
#include <mutex>

#define TDF_THREADS 256
#define TDF_TILE 1024   // triples per shared-memory tile (12 KB)

// `extra`: also write voxel index dim^3 when it is < 27000 — the reference's bound check is `>` not `>=`
// (kernel.cu:13), so thread dim^3 computes and stores one more element inside the 27000-float buffer.
__global__ void __launch_bounds__(TDF_THREADS) k_tdf_batch(const int* __restrict__ occ, const int* __restrict__ occ_begin,
                                                           const int* __restrict__ occ_end, int dim,
                                                           int nvox_write, size_t out_stride, float* __restrict__ out) {
    __shared__ int tile[TDF_TILE * 3];
    int g = blockIdx.y;
    int o0 = occ_begin[g], o1 = occ_end[g];
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    int d2 = dim * dim;
    int z = v / d2, y = (v - z * d2) / dim, x = v - z * d2 - y * dim;
    int best = 900;
    for (int base = o0; base < o1; base += TDF_TILE) {
        int cnt = min(TDF_TILE, o1 - base);
        __syncthreads();
        for (int e = threadIdx.x; e < cnt * 3; e += blockDim.x) tile[e] = __ldg(occ + (size_t)base * 3 + e);
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < cnt; ++i) {
            int ddx = x - tile[i * 3 + 0], ddy = y - tile[i * 3 + 1], ddz = z - tile[i * 3 + 2];
            best = min(best, ddx * ddx + ddy * ddy + ddz * ddz);
        }
    }
    if (v < nvox_write) out[(size_t)g * out_stride + v] = (float)best;
}

// gridDim.y carries the keypoint and is limited to 65535: larger batches go in chunks
#define TDF_MAX_GRID_Y 65535
static int tdf_launch_chunks(rtr_context* ctx, const int* d_occ, const int* d_begin, const int* d_end, int n_grids, int dim, int nvox_write,
                             size_t out_stride, float* d_out) {
    for (int g0 = 0; g0 < n_grids; g0 += TDF_MAX_GRID_Y) {
        int ng = n_grids - g0 < TDF_MAX_GRID_Y ? n_grids - g0 : TDF_MAX_GRID_Y;
        dim3 grid((nvox_write + TDF_THREADS - 1) / TDF_THREADS, ng);
        k_tdf_batch<<<grid, TDF_THREADS, 0, ctx->stream>>>(d_occ, d_begin + g0, d_end + g0, dim, nvox_write, out_stride, d_out + (size_t)g0 * out_stride);
        RTR_LAUNCH_CHECK(ctx, "tdf");
    }
    return 0;
}
static int tdf_launch(rtr_context* ctx, const int* d_occ, const int* d_off, int n_grids, int dim, int nvox_write, size_t out_stride, float* d_out) {
    return tdf_launch_chunks(ctx, d_occ, d_off, d_off + 1, n_grids, dim, nvox_write, out_stride, d_out);
}

// per-grid [begin, end) triple ranges instead of a prefix array (native.cu: fixed-stride voxel lists per keypoint);
// output stride 27000 floats = KeyPoint::grid_value
int rtr_tdf_launch_ranges(rtr_context* ctx, const int* d_occ, const int* d_begin, const int* d_end, int n_grids, int dim, float* d_out) {
    int nv = dim * dim * dim;
    return tdf_launch_chunks(ctx, d_occ, d_begin, d_end, n_grids, dim, nv, (size_t)RTR_TDF_VOXELS, d_out);
}

// process-wide state behind the legacy entry point: device 0 (kernel.cu:43), persistent buffers, pinned staging
static std::mutex g_legacy_mu;
static rtr_context* g_legacy_ctx = nullptr;
static int* g_legacy_occ = nullptr; static int g_legacy_occ_cap = 0;
static int* g_legacy_off = nullptr;
static float* g_legacy_tdf = nullptr;
static int* g_legacy_pin_occ = nullptr; static int g_legacy_pin_cap = 0;
static float* g_legacy_pin_tdf = nullptr;

extern "C" {

int rtr_tdf_batch_dev(rtr_context* ctx, const int* dev_occ, const int* dev_occ_offsets, int n_grids, int dim, float* dev_tdf_out) {
    if (!ctx || n_grids < 0 || dim < 1 || dim > RTR_TDF_DIM || (n_grids > 0 && (!dev_occ_offsets || !dev_tdf_out)))
        return rtr_fail("tdf", "bad argument", RTR_ERR_INVALID);
    RTR_CHECK(cudaSetDevice(ctx->device), "tdf");
    int nv = dim * dim * dim;
    return tdf_launch(ctx, dev_occ, dev_occ_offsets, n_grids, dim, nv, (size_t)nv, dev_tdf_out);
}

int rtr_tdf_batch(rtr_context* ctx, const int* host_occ, const int* host_occ_offsets, int n_grids, int dim, float* host_tdf_out) {
    if (!ctx || n_grids < 0 || dim < 1 || dim > RTR_TDF_DIM || (n_grids > 0 && (!host_occ_offsets || !host_tdf_out)))
        return rtr_fail("tdf", "bad argument", RTR_ERR_INVALID);
    if (n_grids == 0) return 0;
    for (int g = 0; g < n_grids; ++g) if (host_occ_offsets[g + 1] < host_occ_offsets[g] || host_occ_offsets[0] != 0)
        return rtr_fail("tdf", "occ_offsets must start at 0 and be non-decreasing", RTR_ERR_INVALID);
    int total = host_occ_offsets[n_grids];
    if (total > 0 && !host_occ) return rtr_fail("tdf", "null occupied list", RTR_ERR_INVALID);
    RTR_CHECK(cudaSetDevice(ctx->device), "tdf");
    TmpScope tmp_scope(ctx);
    int nv = dim * dim * dim;
    int *d_occ = nullptr, *d_off = nullptr; float* d_out = nullptr;
    if (int e = tmp_alloc(ctx, &d_occ, (size_t)total * 3, "tdf")) return e;
    if (int e = tmp_alloc(ctx, &d_off, (size_t)n_grids + 1, "tdf")) return e;
    if (int e = tmp_alloc(ctx, &d_out, (size_t)n_grids * nv, "tdf")) return e;
    if (total > 0) RTR_CHECK(cudaMemcpyAsync(d_occ, host_occ, (size_t)total * 12, cudaMemcpyHostToDevice, ctx->stream), "tdf.h2d");
    RTR_CHECK(cudaMemcpyAsync(d_off, host_occ_offsets, ((size_t)n_grids + 1) * 4, cudaMemcpyHostToDevice, ctx->stream), "tdf.h2d");
    if (int e = tdf_launch(ctx, d_occ, d_off, n_grids, dim, nv, (size_t)nv, d_out)) return e;
    RTR_CHECK(cudaMemcpyAsync(host_tdf_out, d_out, (size_t)n_grids * nv * 4, cudaMemcpyDeviceToHost, ctx->stream), "tdf.d2h");
    RTR_CHECK(cudaStreamSynchronize(ctx->stream), "tdf");
    dev_free(ctx, d_occ); dev_free(ctx, d_off); dev_free(ctx, d_out);
    return 0;
}

// The reference FFI, same symbol, same signature, same semantics (host pointers, synchronous, 0 on success).
int ComputeTDFWithCuda(const int* voxel_grid_occ, float* voxel_grid_TDF, int voxel_grid_dim, int num_occ) {
    std::lock_guard<std::mutex> lock(g_legacy_mu);
    if (!voxel_grid_TDF || voxel_grid_dim < 1 || voxel_grid_dim > RTR_TDF_DIM) {
        fprintf(stderr, "rtr[ComputeTDFWithCuda] invalid grid (dim must be 1..30, buffer 27000 floats)\n");
        return RTR_ERR_INVALID;
    }
    if (num_occ < 0 || (num_occ > 0 && !voxel_grid_occ)) {
        // key_point.h:296,313 passes num_center - 1 == -1 when the keypoint has no occupied voxel; the reference then
        // fails inside cudaMalloc.  Report it the same way (non-zero status, one stderr line) instead of crashing.
        fprintf(stderr, "rtr[ComputeTDFWithCuda] invalid occupied-voxel list (num_occ = %d)\n", num_occ);
        return RTR_ERR_INVALID;
    }
    if (!g_legacy_ctx) {
        // the context is published only once every persistent buffer exists: a failed first call leaves nothing half
        // initialised behind (the next call starts over)
        rtr_context* c0 = nullptr;
        float *tdf0 = nullptr, *pin0 = nullptr; int* off0 = nullptr;
        int e = rtr_context_create(0, &c0);
        if (!e) e = (int)cudaMalloc(&tdf0, RTR_TDF_VOXELS * sizeof(float));
        if (!e) e = (int)cudaMalloc(&off0, 2 * sizeof(int));
        if (!e) e = (int)cudaMallocHost(&pin0, RTR_TDF_VOXELS * sizeof(float));
        if (e) {
            fprintf(stderr, "rtr[ComputeTDFWithCuda] initialisation failed (status %d)\n", e);
            if (pin0) cudaFreeHost(pin0);
            if (off0) cudaFree(off0);
            if (tdf0) cudaFree(tdf0);
            if (c0) rtr_context_destroy(c0);
            return e;
        }
        g_legacy_tdf = tdf0; g_legacy_off = off0; g_legacy_pin_tdf = pin0; g_legacy_ctx = c0;
    }
    rtr_context* ctx = g_legacy_ctx;
    RTR_CHECK(cudaSetDevice(0), "ComputeTDFWithCuda");
    if (num_occ > g_legacy_occ_cap) {
        if (g_legacy_occ) cudaFree(g_legacy_occ);
        int cap = num_occ < 1024 ? 1024 : num_occ * 2;
        RTR_CHECK(cudaMalloc(&g_legacy_occ, (size_t)cap * 12), "ComputeTDFWithCuda");
        g_legacy_occ_cap = cap;
    }
    if (num_occ + 1 > g_legacy_pin_cap) {
        if (g_legacy_pin_occ) cudaFreeHost(g_legacy_pin_occ);
        int cap = num_occ < 1024 ? 1025 : num_occ * 2 + 1;
        RTR_CHECK(cudaMallocHost(&g_legacy_pin_occ, (size_t)cap * 12 + 8), "ComputeTDFWithCuda");
        g_legacy_pin_cap = cap;
    }
    int nv = voxel_grid_dim * voxel_grid_dim * voxel_grid_dim;
    int nw = nv < RTR_TDF_VOXELS ? nv + 1 : nv;     // kernel.cu:13 `>`: element dim^3 is also written when it exists
    // offsets {0, num_occ} ride in front of the triples in one pinned buffer -> a single H2D for the inputs
    int* pin = g_legacy_pin_occ;
    pin[0] = 0; pin[1] = num_occ;
    if (num_occ > 0) memcpy(pin + 2, voxel_grid_occ, (size_t)num_occ * 12);
    RTR_CHECK(cudaMemcpyAsync(g_legacy_off, pin, 8, cudaMemcpyHostToDevice, ctx->stream), "ComputeTDFWithCuda");
    if (num_occ > 0) RTR_CHECK(cudaMemcpyAsync(g_legacy_occ, pin + 2, (size_t)num_occ * 12, cudaMemcpyHostToDevice, ctx->stream), "ComputeTDFWithCuda");
    if (int e = tdf_launch(ctx, g_legacy_occ, g_legacy_off, 1, voxel_grid_dim, nw, (size_t)RTR_TDF_VOXELS, g_legacy_tdf)) return e;
    RTR_CHECK(cudaMemcpyAsync(g_legacy_pin_tdf, g_legacy_tdf, (size_t)nw * 4, cudaMemcpyDeviceToHost, ctx->stream), "ComputeTDFWithCuda");
    RTR_CHECK(cudaStreamSynchronize(ctx->stream), "ComputeTDFWithCuda");
    // elements [nw, 27000) keep the caller's values: the reference copies the caller's buffer in and back out
    memcpy(voxel_grid_TDF, g_legacy_pin_tdf, (size_t)nw * 4);
    return 0;
}

}  // extern "C"
