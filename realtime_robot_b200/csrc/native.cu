// native.cu — the reference's own descriptor path, batched on the device:
//   KeyPoint::getOccupiedGrid  key_point.h:112-161     k_native_occupancy   one CTA per keypoint
//   KeyPoint::get_TSDF         key_point.h:251-318     k_native_tdf_voxels  one CTA per keypoint -> k_tdf_batch (tdf.cu)
//   get_Distance               matching.h:122-222      k_native_pair_score  one CTA per (model keypoint, scan keypoint):
//                                                       the model TDF (108 KB) lives in shared memory for all 36 angles
//   match_by_* + pair loop + Ransac   function.h:158-178, RealTimeRobot.cpp:70-102, function.h:35-109   k_native_consensus
//
// The reference rebuilds a whole-cloud octree per keypoint, runs one synchronous TDF launch per keypoint and calls
// get_Distance Km*Ks + P*(P+1) times (each building 36 octrees).  Here every keypoint / pair is one CTA of one launch,
// voxel sets are 32^3-bit bitmaps in shared memory kept in the octree's depth-first (Morton, x most significant) order,
// and pair scores are computed once and reused by the consensus step.
//
// Semantics: the INTENDED algorithm with the quirk_* flags of rtr_native_params (documented with the test restatement and in
// DESIGN.md); voxel frames follow pcl::octree::OctreePointCloud in fp64, identical to the oracle.
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>

int rtr_tdf_launch_ranges(rtr_context* ctx, const int* d_occ, const int* d_begin, const int* d_end, int n_grids, int dim, float* d_out);

#define NAT_THREADS 256
#define NAT_WORDS 1024          // 32^3 bits
#define NAT_LIST 4096           // occupied voxels of one sweep angle that fit the shared-memory list

struct NatFrame { double origin[3]; double mn[3]; double res; int depth; unsigned nkeys; };

// OctreePointCloud::defineBoundingBox + getKeyBitSize for the float box kp -/+ half
__device__ __forceinline__ NatFrame nat_frame(float4 kp, float half, float resolution) {
    NatFrame f;
    f.res = (double)resolution;
    const float c[3] = {kp.x, kp.y, kp.z};
    double mx[3];
    unsigned max_voxels = 2;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        f.mn[a] = (double)__fsub_rn(c[a], half); mx[a] = (double)__fadd_rn(c[a], half);
        max_voxels = max(max_voxels, (unsigned)ceil((mx[a] - f.mn[a] - (double)FLT_EPSILON) / f.res));
    }
    f.depth = max((int)ceil(log2((double)max_voxels) - (double)FLT_EPSILON), 0);
    f.depth = min(f.depth, 5);                      // bitmaps hold 32 keys per axis (the reference's boxes give exactly 5)
    f.nkeys = 1u << f.depth;
    double side = (double)f.nkeys * f.res;
#pragma unroll
    for (int a = 0; a < 3; ++a) f.origin[a] = f.mn[a] - (side - (mx[a] - f.mn[a])) / 2.0;
    return f;
}
__device__ __forceinline__ bool nat_key(const NatFrame& f, float4 p, unsigned& code) {
    const double v[3] = {p.x, p.y, p.z};
    unsigned k[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double q = (v[a] - f.origin[a]) / f.res;
        if (!(q >= 0.0) || q >= (double)f.nkeys) return false;
        k[a] = (unsigned)q;
    }
    code = 0;
    for (int b = f.depth - 1; b >= 0; --b) code = (code << 3) | (((k[0] >> b) & 1u) << 2) | (((k[1] >> b) & 1u) << 1) | ((k[2] >> b) & 1u);
    return true;
}
__device__ __forceinline__ void nat_unmorton(unsigned code, int depth, unsigned k[3]) {
    k[0] = k[1] = k[2] = 0;
    for (int b = 0; b < depth; ++b) {
        unsigned t = (code >> (3 * b)) & 7u;
        k[0] |= ((t >> 2) & 1u) << b; k[1] |= ((t >> 1) & 1u) << b; k[2] |= (t & 1u) << b;
    }
}
// truncating cast of (voxel centre - box minimum) / resolution (key_point.h:300-307, matching.h:181-183)
__device__ __forceinline__ int nat_index(const NatFrame& f, unsigned k, int a, float resolution) {
    float centre = (float)(((double)k + 0.5) * f.res + f.origin[a]);
    return (int)(((double)centre - f.mn[a]) / (double)resolution);
}

// ---- KeyPoint::getOccupiedGrid: points in the inclusive +-half box and the number of occupied voxels
__global__ void __launch_bounds__(NAT_THREADS) k_native_occupancy(GridView g, const float4* __restrict__ kps, int n_kp, float half,
                                                                  float resolution, int cap, float4* __restrict__ occ_pts,
                                                                  int* __restrict__ occ_count, int* __restrict__ number) {
    __shared__ unsigned bits[NAT_WORDS];
    __shared__ int s_count, s_number;
    int k = blockIdx.x;
    if (k >= n_kp) return;
    for (int w = threadIdx.x; w < NAT_WORDS; w += NAT_THREADS) bits[w] = 0u;
    if (threadIdx.x == 0) { s_count = 0; s_number = 0; }
    __syncthreads();
    float4 kp = kps[k];
    float lo[3] = {__fsub_rn(kp.x, half), __fsub_rn(kp.y, half), __fsub_rn(kp.z, half)};
    float hi[3] = {__fadd_rn(kp.x, half), __fadd_rn(kp.y, half), __fadd_rn(kp.z, half)};
    NatFrame f = nat_frame(kp, half, resolution);
    int cx = cell_coord(kp.x, g.mnx, g.inv_h), cy = cell_coord(kp.y, g.mny, g.inv_h), cz = cell_coord(kp.z, g.mnz, g.inv_h);
    if (!(cx < -1 || cy < -1 || cz < -1 || cx > g.dx || cy > g.dy || cz > g.dz)) {
        cx = clampi(cx, 0, g.dx - 1); cy = clampi(cy, 0, g.dy - 1); cz = clampi(cz, 0, g.dz - 1);
        int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dx - 1);
        for (int z = max(cz - 1, 0); z <= min(cz + 1, g.dz - 1); ++z)
            for (int y = max(cy - 1, 0); y <= min(cy + 1, g.dy - 1); ++y) {
                int s0 = __ldg(g.cell_begin + cell_key(g, x0, y, z));
                int s1 = __ldg(g.cell_begin + cell_key(g, x1, y, z) + 1);
                for (int s = s0 + threadIdx.x; s < s1; s += NAT_THREADS) {
                    float4 p = __ldg(g.sorted + s);
                    if (p.x >= lo[0] && p.x <= hi[0] && p.y >= lo[1] && p.y <= hi[1] && p.z >= lo[2] && p.z <= hi[2]) {
                        int slot = atomicAdd(&s_count, 1);
                        p.w = 1.0f;
                        if (slot < cap) occ_pts[(size_t)k * cap + slot] = p;
                        unsigned code;
                        if (nat_key(f, p, code)) atomicOr(&bits[code >> 5], 1u << (code & 31u));
                    }
                }
            }
    }
    __syncthreads();
    int c = 0;
    for (int w = threadIdx.x; w < NAT_WORDS; w += NAT_THREADS) c += __popc(bits[w]);
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_number, c);
    __syncthreads();
    if (threadIdx.x == 0) { occ_count[k] = s_count; number[k] = s_number; }
}

// ordered enumeration of the set bits of a Morton-ordered bitmap: every thread owns 4 consecutive words; returns the
// exclusive prefix of its set-bit count and the block total (scan through shared memory)
__device__ __forceinline__ int nat_prefix(const unsigned* bits, int* scan /*[NAT_THREADS]*/, int& total) {
    int t = threadIdx.x, c = 0;
#pragma unroll
    for (int w = 0; w < NAT_WORDS / NAT_THREADS; ++w) c += __popc(bits[t * (NAT_WORDS / NAT_THREADS) + w]);
    scan[t] = c;
    __syncthreads();
    for (int off = 1; off < NAT_THREADS; off <<= 1) {
        int v = (t >= off) ? scan[t - off] : 0;
        __syncthreads();
        scan[t] += v;
        __syncthreads();
    }
    total = scan[NAT_THREADS - 1];
    return scan[t] - c;
}

// ---- KeyPoint::get_TSDF up to the TDF call: occupied voxels of the occupancy cloud in the +-tdf_half frame, in
// depth-first order, as int triples relative to the box minimum
__global__ void __launch_bounds__(NAT_THREADS) k_native_tdf_voxels(const float4* __restrict__ kps, int n_kp, float half, float resolution,
                                                                   int cap, const float4* __restrict__ occ_pts, const int* __restrict__ occ_count,
                                                                   int skip_first, int vcap, int* __restrict__ triples,
                                                                   int* __restrict__ v_begin, int* __restrict__ v_end) {
    __shared__ unsigned bits[NAT_WORDS];
    __shared__ int scan[NAT_THREADS];
    int k = blockIdx.x;
    if (k >= n_kp) return;
    for (int w = threadIdx.x; w < NAT_WORDS; w += NAT_THREADS) bits[w] = 0u;
    __syncthreads();
    float4 kp = kps[k];
    NatFrame f = nat_frame(kp, half, resolution);
    int m = min(occ_count[k], cap);
    for (int i = threadIdx.x; i < m; i += NAT_THREADS) {
        unsigned code;
        if (nat_key(f, occ_pts[(size_t)k * cap + i], code)) atomicOr(&bits[code >> 5], 1u << (code & 31u));
    }
    __syncthreads();
    int total;
    int pos = nat_prefix(bits, scan, total);
    const int wpt = NAT_WORDS / NAT_THREADS;
    for (int w = 0; w < wpt; ++w) {
        unsigned word = bits[threadIdx.x * wpt + w];
        while (word) {
            int b = __ffs(word) - 1;
            word &= word - 1;
            int out = pos - (skip_first ? 1 : 0);
            if (out >= 0 && out < vcap) {
                unsigned kk[3];
                nat_unmorton((unsigned)((threadIdx.x * wpt + w) * 32 + b), f.depth, kk);
                int* t = triples + ((size_t)k * vcap + out) * 3;
#pragma unroll
                for (int a = 0; a < 3; ++a) t[a] = nat_index(f, kk[a], a, resolution);
            }
            ++pos;
        }
    }
    if (threadIdx.x == 0) {
        int cnt = max(total - (skip_first ? 1 : 0), 0);
        v_begin[k] = k * vcap;
        v_end[k] = k * vcap + min(cnt, vcap);
    }
}

// ---- get_Distance for one (model keypoint, scan keypoint) pair per CTA
// cos / sin come from the HOST (glibc, like the oracle and like the reference's CPU code): the per-step rotation uses
// cos(float theta) (matching.h:152-156), the final rotation cos(double(i * theta)) (matching.h:195,205-208)
struct NatSweep { float step_c, step_s; float cs[36][2]; };

__global__ void __launch_bounds__(NAT_THREADS) k_native_pair_score(const float4* __restrict__ mkp, int km, const float4* __restrict__ skp, int ks,
                                                                   const float* __restrict__ mtdf, const float4* __restrict__ socc,
                                                                   const int* __restrict__ socc_count, int cap, float4* __restrict__ scratch,
                                                                   NatSweep sw, float half, float resolution, int skip_first, int running,
                                                                   float* __restrict__ score, int* __restrict__ best_step, float* __restrict__ transform) {
    extern __shared__ __align__(16) unsigned char nat_smem[];
    float* tdf = reinterpret_cast<float*>(nat_smem);                         // 27000 floats
    unsigned* bits = reinterpret_cast<unsigned*>(nat_smem + 27000 * 4);     // 1024 words
    float* list = reinterpret_cast<float*>(nat_smem + 27000 * 4 + NAT_WORDS * 4);   // NAT_LIST floats
    __shared__ float s_step[16];
    __shared__ int scan[NAT_THREADS];
    int pair = blockIdx.x;
    int k = pair / ks, s = pair - k * ks;
    if (k >= km) return;
    for (int i = threadIdx.x; i < 27000; i += NAT_THREADS) tdf[i] = mtdf[(size_t)k * 27000 + i];
    float4 k1 = mkp[k], k2 = skp[s];
    if (threadIdx.x == 0) {
        // transform_3 * transform_2 * transform_1 (matching.h:147-167): rotation by 10 degrees about (k1.x, k1.y)
        float T1[16], R[16], T3[16], tmp[16];
        for (int i = 0; i < 16; ++i) { T1[i] = T3[i] = R[i] = (i % 5 == 0) ? 1.f : 0.f; }
        T1[12] = -k1.x; T1[13] = -k1.y; T3[12] = k1.x; T3[13] = k1.y;
        R[0] = sw.step_c; R[1] = sw.step_s; R[4] = -sw.step_s; R[5] = sw.step_c;
        matmul4(T3, R, tmp); matmul4(tmp, T1, tmp);
        for (int i = 0; i < 16; ++i) s_step[i] = tmp[i];
    }
    // translate the scan keypoint's occupancy cloud so that the keypoints coincide (matching.h:130-138)
    float T[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.f : 0.f;
    T[12] = __fsub_rn(k1.x, k2.x); T[13] = __fsub_rn(k1.y, k2.y); T[14] = __fsub_rn(k1.z, k2.z);
    int m = min(socc_count[s], cap);
    float4* cloud = scratch + (size_t)pair * cap;
    for (int i = threadIdx.x; i < m; i += NAT_THREADS) cloud[i] = xform(T, socc[(size_t)s * cap + i]);
    NatFrame f = nat_frame(k1, half, resolution);
    float distance_total = 100000000.f, distance_temp = 0.f;
    int best = 0;
    __syncthreads();
    for (int ang = 0; ang < 36; ++ang) {
        for (int w = threadIdx.x; w < NAT_WORDS; w += NAT_THREADS) bits[w] = 0u;
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += NAT_THREADS) {
            float4 p = cloud[i];
            if (ang != 0) { p = xform(s_step, p); cloud[i] = p; }            // cumulative, in place, float (Appendix B#13)
            unsigned code;
            if (nat_key(f, p, code)) atomicOr(&bits[code >> 5], 1u << (code & 31u));
        }
        __syncthreads();
        // All threads enumerate their share of the bitmap (ordered by an exclusive scan of the per-thread bit counts) and
        // write grid_value^2 of every occupied voxel into a list in depth-first order (-1 marks voxels outside the TDF);
        // thread 0 then does what must stay sequential: the float accumulation in that order (matching.h:179-189).
        int total;
        int pos = nat_prefix(bits, scan, total);
        const bool fits = total <= NAT_LIST;
        if (fits) {
            const int wpt = NAT_WORDS / NAT_THREADS;
            for (int w = 0; w < wpt; ++w) {
                unsigned word = bits[threadIdx.x * wpt + w];
                while (word) {
                    int b = __ffs(word) - 1;
                    word &= word - 1;
                    unsigned kk[3];
                    nat_unmorton((unsigned)((threadIdx.x * wpt + w) * 32 + b), f.depth, kk);
                    int cx = nat_index(f, kk[0], 0, resolution), cy = nat_index(f, kk[1], 1, resolution), cz = nat_index(f, kk[2], 2, resolution);
                    float v = -1.f;
                    if (cx >= 0 && cy >= 0 && cz >= 0 && cx < 30 && cy < 30 && cz < 30) { float gv = tdf[cy * 30 + cz * 900 + cx]; v = __fmul_rn(gv, gv); }
                    list[pos++] = v;
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (!running) distance_temp = 0.f;
            int used = 0;
            if (fits) {
                for (int i = skip_first ? 1 : 0; i < total; ++i) {
                    float v = list[i];
                    if (v >= 0.f) { distance_temp = __fadd_rn(distance_temp, v); ++used; }
                }
            } else {                     // more occupied voxels than the list holds: serial enumeration
                int seen = 0;
                for (int w = 0; w < NAT_WORDS; ++w) {
                    unsigned word = bits[w];
                    while (word) {
                        int b = __ffs(word) - 1;
                        word &= word - 1;
                        if (!(skip_first && seen == 0)) {
                            unsigned kk[3];
                            nat_unmorton((unsigned)(w * 32 + b), f.depth, kk);
                            int cx = nat_index(f, kk[0], 0, resolution), cy = nat_index(f, kk[1], 1, resolution), cz = nat_index(f, kk[2], 2, resolution);
                            if (cx >= 0 && cy >= 0 && cz >= 0 && cx < 30 && cy < 30 && cz < 30) {
                                float gv = tdf[cy * 30 + cz * 900 + cx];
                                distance_temp = __fadd_rn(distance_temp, __fmul_rn(gv, gv));
                                ++used;
                            }
                        }
                        ++seen;
                    }
                }
            }
            distance_temp = used > 0 ? __fdiv_rn(distance_temp, (float)used) : 100000000.f;
            if (distance_temp < distance_total) { distance_total = distance_temp; best = ang; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        score[pair] = distance_total;
        best_step[pair] = best;
        // key_transform = S * T3 * R(best) * T1 * T (matching.h:204-217)
        float T1[16], R2[16], T3[16], S[16], out[16];
        for (int i = 0; i < 16; ++i) { T1[i] = T3[i] = R2[i] = S[i] = (i % 5 == 0) ? 1.f : 0.f; }
        T1[12] = -k1.x; T1[13] = -k1.y; T3[12] = k1.x; T3[13] = k1.y;
        R2[0] = sw.cs[best][0]; R2[1] = sw.cs[best][1]; R2[4] = -sw.cs[best][1]; R2[5] = sw.cs[best][0];
        float cs = __fdiv_rn(k1.z, k2.z);
        S[0] = S[5] = S[10] = cs;
        matmul4(S, T3, out); matmul4(out, R2, out); matmul4(out, T1, out); matmul4(out, T, out);
        for (int i = 0; i < 16; ++i) transform[(size_t)pair * 16 + i] = out[i];
    }
}

// KeyPoint::get_Vector3D (key_point.h:87-111) over the kept surfaces of a cloud, with getDistance (key_point.h:38-43)
struct NatSurface { float c[4]; double area; int is_vertical; int pad; };
__device__ void nat_vector3d(float4 kp, const NatSurface* __restrict__ surf, int ns, double* v3) {
    v3[0] = v3[1] = v3[2] = 0.16;
    int vertical = 0, horizontal = 0;
    for (int i = 0; i < ns; ++i) {
        const float* v = surf[i].c;
        double d = (double)__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(v[0], v[0]), __fmul_rn(v[1], v[1])), __fmul_rn(v[2], v[2])));
        double distance = (double)fabsf(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v[0], kp.x), __fmul_rn(v[1], kp.y)), __fmul_rn(v[2], kp.z)), v[3])) / d;
        if (surf[i].is_vertical == 0 && horizontal == 0 && distance <= 0.05) { v3[0] = surf[i].area; horizontal++; }
        else if (surf[i].is_vertical == 1 && vertical <= 1 && distance <= 0.02) { v3[1 + vertical] = surf[i].area; vertical++; }
    }
    if (v3[1] < v3[2]) { double t = v3[1]; v3[1] = v3[2]; v3[2] = t; }
}

// ---- screens (function.h:158-178), pair list (RealTimeRobot.cpp:70-102) and exhaustive consensus (function.h:35-109)
__global__ void __launch_bounds__(1024) k_native_consensus(const float4* __restrict__ mkp, int km, const float4* __restrict__ skp, int ks,
                                                           const int* __restrict__ mnum, const int* __restrict__ snum,
                                                           const float* __restrict__ score, const float* __restrict__ transform,
                                                           const NatSurface* __restrict__ msurf, int n_msurf, const NatSurface* __restrict__ ssurf,
                                                           int n_ssurf, rtr_native_params p, int* __restrict__ pair_list,
                                                           rtr_pose_result* __restrict__ res) {
    __shared__ int s_np;
    __shared__ int s_in[1024];
    int npairs = km * ks;
    if (threadIdx.x == 0) {
        // ordered compaction of the pairs that pass the gate and the screens (the areas are the KeyPoint defaults 0.16)
        int c = 0;
        for (int q = 0; q < npairs; ++q) {
            int k = q / ks, s = q - k * ks;
            bool gate = score[q] < p.pair_gate;
            float hr = __fdiv_rn(mkp[k].z, skp[s].z);
            bool height = p.quirk_integer_screens ? (hr >= 0.0f || hr <= 1.5f) : (hr >= 2.0f / 3.0f && hr <= 1.5f);
            bool occ;
            if (snum[s] == 0) occ = false;
            else {
                float t = p.quirk_integer_screens ? (float)(mnum[k] / snum[s]) : __fdiv_rn((float)mnum[k], (float)snum[s]);
                occ = !(t > 2.f || t < 0.5f);
            }
            // match_by_area (function.h:165-170) on the keypoints' three plane areas
            bool area = true;
            if (gate && height && occ) {
                double a1[3], a2[3];
                nat_vector3d(mkp[k], msurf, n_msurf, a1);
                nat_vector3d(skp[s], ssurf, n_ssurf, a2);
                double lo = p.quirk_integer_screens ? 0.0 : 1.0 / 3.0;
                for (int a = 0; a < 3; ++a) { double r = a1[a] / a2[a]; if (r > 3 || r < lo) area = false; }
            }
            if (gate && height && occ && area) pair_list[c++] = q;
        }
        s_np = c;
        for (int i = 0; i < 16; ++i) res->pose[i] = (i % 5 == 0) ? 1.f : 0.f;
        res->fitness = FLT_MAX; res->inliers = 0; res->hypothesis = -1; res->evaluated = c; res->converged = 0; res->iterations = 0;
        res->model_id = 0; res->n_keypoints_src = km; res->n_keypoints_tgt = ks;
        for (int i = 0; i < 5; ++i) res->pad_[i] = 0;
    }
    __syncthreads();
    int P = s_np;
    int best_in = 0, best_i = -1;
    for (int base = 0; base < P; base += blockDim.x) {
        int i = base + threadIdx.x;
        int in = 0;
        if (i < P) {
            const float* T = transform + (size_t)pair_list[i] * 16;
            for (int j = 0; j < P; ++j) {
                int q = pair_list[j];
                float4 a = mkp[q / ks];
                float4 b = xform(T, skp[q % ks]);
                float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y), dz = __fsub_rn(a.z, b.z);
                float sumf = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                double dis = (double)__fsqrt_rn(sumf);                    // pointdistance (function.h:27-30)
                if (dis < (double)p.consensus_distance && score[q] < p.consensus_score) ++in;
            }
        }
        s_in[threadIdx.x] = in;
        __syncthreads();
        if (threadIdx.x == 0)
            for (int t = 0; t < (int)blockDim.x && base + t < P; ++t)
                if (s_in[t] > best_in) { best_in = s_in[t]; best_i = base + t; }     // first arg-max (function.h:83-87)
        __syncthreads();
    }
    if (threadIdx.x == 0 && best_i >= 0) {
        int q = pair_list[best_i];
        for (int i = 0; i < 16; ++i) res->pose[i] = transform[(size_t)q * 16 + i];
        res->inliers = best_in; res->hypothesis = best_i; res->fitness = score[q]; res->converged = 1;
    }
}

// ----------------------------------------------------------------------------- host
struct NatDesc {              // per-cloud keypoint descriptors on the device
    int n_kp = 0, cap = 0, vcap = 0;
    float4* kps = nullptr; float4* occ = nullptr; int* occ_count = nullptr; int* number = nullptr;
    int* triples = nullptr; int* v_begin = nullptr; int* v_end = nullptr; float* tdf = nullptr;
};
static void nat_free(rtr_context* ctx, NatDesc& d) {
    dev_free(ctx, d.kps); dev_free(ctx, d.occ); dev_free(ctx, d.occ_count); dev_free(ctx, d.number);
    dev_free(ctx, d.triples); dev_free(ctx, d.v_begin); dev_free(ctx, d.v_end); dev_free(ctx, d.tdf);
    d = NatDesc();
}

// d_kps: device keypoints (n_kp x float4).  with_tdf: also the TDF (model keypoints only, RealTimeRobot.cpp:56 vs :66)
static int nat_describe(rtr_cloud* c, const float4* d_kps, int n_kp, const rtr_native_params* p, bool with_tdf, NatDesc& d) {
    rtr_context* ctx = c->ctx;
    d.n_kp = n_kp;
    d.cap = std::max(1, std::min(c->n, 16384));
    d.vcap = std::min(d.cap, 32768);
    if (int e = tmp_alloc(ctx, &d.kps, n_kp, "native")) return e;
    if (int e = tmp_alloc(ctx, &d.occ, (size_t)n_kp * d.cap, "native")) return e;
    if (int e = tmp_alloc(ctx, &d.occ_count, n_kp, "native")) return e;
    if (int e = tmp_alloc(ctx, &d.number, n_kp, "native")) return e;
    if (n_kp == 0) return 0;
    RTR_CHECK(cudaMemcpyAsync(d.kps, d_kps, (size_t)n_kp * 16, cudaMemcpyDeviceToDevice, ctx->stream), "native");
    DevGrid* g;
    if (int e = rtr_get_grid(c, p->occ_half, &g)) return e;
    k_native_occupancy<<<n_kp, NAT_THREADS, 0, ctx->stream>>>(rtr_view(g), d.kps, n_kp, p->occ_half, p->resolution, d.cap, d.occ, d.occ_count, d.number);
    RTR_LAUNCH_CHECK(ctx, "native.occupancy");
    if (with_tdf) {
        if (int e = tmp_alloc(ctx, &d.triples, (size_t)n_kp * d.vcap * 3, "native")) return e;
        if (int e = tmp_alloc(ctx, &d.v_begin, n_kp, "native")) return e;
        if (int e = tmp_alloc(ctx, &d.v_end, n_kp, "native")) return e;
        if (int e = tmp_alloc(ctx, &d.tdf, (size_t)n_kp * 27000, "native")) return e;
        k_native_tdf_voxels<<<n_kp, NAT_THREADS, 0, ctx->stream>>>(d.kps, n_kp, p->tdf_half, p->resolution, d.cap, d.occ, d.occ_count,
                                                                    p->quirk_skip_first_voxel, d.vcap, d.triples, d.v_begin, d.v_end);
        RTR_LAUNCH_CHECK(ctx, "native.tdf_voxels");
        // dim = f_adjust / resolution * 2 in float (key_point.h:310)
        int dim = (int)(p->tdf_half / p->resolution * 2);
        if (dim < 1 || dim > RTR_TDF_DIM) return rtr_fail("native", "tdf_half / resolution * 2 must be in 1..30", RTR_ERR_INVALID);
        if (int e = rtr_tdf_launch_ranges(ctx, d.triples, d.v_begin, d.v_end, n_kp, dim, d.tdf)) return e;
    }
    return 0;
}

static NatSweep nat_sweep() {
    NatSweep sw;
    float theta = (float)(M_PI / 18);                      // matching.h:143
    sw.step_c = std::cos(theta); sw.step_s = std::sin(theta);          // float overloads, as in the reference
    for (int i = 0; i < 36; ++i) {
        double bt = (double)i * (double)theta;
        sw.cs[i][0] = (float)std::cos(bt); sw.cs[i][1] = (float)std::sin(bt);
    }
    return sw;
}

static int nat_pairs(rtr_context* ctx, const NatDesc& dm, const NatDesc& ds, const rtr_native_params* p, float** score, int** best,
                     float** transform) {
    int npairs = dm.n_kp * ds.n_kp;
    if (int e = dev_alloc(ctx, score, npairs, "native")) return e;
    if (int e = dev_alloc(ctx, best, npairs, "native")) return e;
    if (int e = dev_alloc(ctx, transform, (size_t)npairs * 16, "native")) return e;
    if (npairs == 0) return 0;
    float4* scratch = nullptr;
    if (int e = tmp_alloc(ctx, &scratch, (size_t)npairs * ds.cap, "native")) return e;
    size_t smem = 27000 * 4 + NAT_WORDS * 4 + NAT_LIST * 4;
    if (int e = rtr_kernel_smem(k_native_pair_score, ctx, smem)) return e;
    NatSweep final_sw = nat_sweep();
    k_native_pair_score<<<npairs, NAT_THREADS, smem, ctx->stream>>>(dm.kps, dm.n_kp, ds.kps, ds.n_kp, dm.tdf, ds.occ, ds.occ_count, ds.cap, scratch,
                                                                     final_sw, p->tdf_half, p->resolution, p->quirk_skip_first_voxel,
                                                                     p->quirk_running_score, *score, *best, *transform);
    RTR_LAUNCH_CHECK(ctx, "native.pair_score");
    dev_free(ctx, scratch);
    return 0;
}

// A zero / negative / NaN half-width or resolution would hang the grid sizing or index outside the TDF; the pair sweep
// (k_native_pair_score) and the consensus are written for the reference's 30^3 TDF (KeyPoint::grid_value[27000]).
static int nat_validate(const rtr_native_params* p, bool need_dim30) {
    if (!(p->resolution > 0.f) || !std::isfinite(p->resolution) || !(p->occ_half > 0.f) || !std::isfinite(p->occ_half) ||
        !(p->tdf_half > 0.f) || !std::isfinite(p->tdf_half))
        return rtr_fail("native", "resolution, occ_half and tdf_half must be finite and > 0", RTR_ERR_INVALID);
    int dim = (int)(p->tdf_half / p->resolution * 2);
    if (dim < 1 || dim > RTR_TDF_DIM) return rtr_fail("native", "tdf_half / resolution * 2 must be in 1..30", RTR_ERR_INVALID);
    if (need_dim30 && dim != RTR_TDF_DIM) return rtr_fail("native", "the pair sweep needs tdf_half / resolution * 2 == 30 (KeyPoint::grid_value[27000])", RTR_ERR_INVALID);
    return 0;
}

extern "C" {

void rtr_native_default_params(rtr_native_params* p) {
    memset(p, 0, sizeof(*p));
    p->resolution = 0.01f; p->occ_half = 0.1f; p->tdf_half = 0.15f;
    p->pair_gate = 3.0f; p->consensus_distance = 0.15f; p->consensus_score = 100.0f;
    p->use_plane_areas = 1;
}

int rtr_native_keypoint_descriptors(rtr_cloud* c, const float* host_kp_xyz1, int n_kp, const rtr_native_params* p, int* host_number,
                                    int* host_occ_count, float* host_tdf, int* host_voxel_count) {
    if (!c || !p || n_kp < 0 || (n_kp > 0 && !host_kp_xyz1)) return rtr_fail("native", "bad argument", RTR_ERR_INVALID);
    if (int e = nat_validate(p, false)) return e;
    rtr_context* ctx = c->ctx;
    TmpScope tmp_scope(ctx);
    RTR_CHECK(cudaSetDevice(ctx->device), "native");
    float4* d_kps = nullptr;
    if (int e = tmp_alloc(ctx, &d_kps, n_kp, "native")) return e;
    if (n_kp) RTR_CHECK(cudaMemcpyAsync(d_kps, host_kp_xyz1, (size_t)n_kp * 16, cudaMemcpyHostToDevice, ctx->stream), "native");
    NatDesc d;
    if (int e = nat_describe(c, d_kps, n_kp, p, host_tdf || host_voxel_count, d)) return e;
    int rc = 0;
    if (n_kp) {
        std::vector<int> cnt(n_kp), vb(n_kp), ve(n_kp);
        RTR_CHECK(cudaMemcpyAsync(cnt.data(), d.occ_count, (size_t)n_kp * 4, cudaMemcpyDeviceToHost, ctx->stream), "native");
        if (host_number) RTR_CHECK(cudaMemcpyAsync(host_number, d.number, (size_t)n_kp * 4, cudaMemcpyDeviceToHost, ctx->stream), "native");
        if (d.tdf) {
            RTR_CHECK(cudaMemcpyAsync(vb.data(), d.v_begin, (size_t)n_kp * 4, cudaMemcpyDeviceToHost, ctx->stream), "native");
            RTR_CHECK(cudaMemcpyAsync(ve.data(), d.v_end, (size_t)n_kp * 4, cudaMemcpyDeviceToHost, ctx->stream), "native");
            if (host_tdf) RTR_CHECK(cudaMemcpyAsync(host_tdf, d.tdf, (size_t)n_kp * 27000 * 4, cudaMemcpyDeviceToHost, ctx->stream), "native");
        }
        RTR_CHECK(cudaStreamSynchronize(ctx->stream), "native");
        for (int k = 0; k < n_kp; ++k) {
            if (cnt[k] > d.cap) rc = rtr_fail("native", "occupancy cloud exceeds the per-keypoint capacity (16384 points)", RTR_ERR_CAPACITY);
            if (host_occ_count) host_occ_count[k] = cnt[k];
            if (host_voxel_count && d.tdf) host_voxel_count[k] = ve[k] - vb[k];
        }
    }
    nat_free(ctx, d); dev_free(ctx, d_kps);
    return rc;
}

int rtr_native_pair_scores(rtr_cloud* model, const float* host_model_kp_xyz1, int km, rtr_cloud* scan, const float* host_scan_kp_xyz1,
                           int ks, const rtr_native_params* p, float* host_score, int* host_best_step, float* host_transform16) {
    if (!model || !scan || !p || km < 0 || ks < 0 || model->ctx != scan->ctx) return rtr_fail("native", "bad argument", RTR_ERR_INVALID);
    if (int e = nat_validate(p, true)) return e;
    rtr_context* ctx = model->ctx;
    TmpScope tmp_scope(ctx);
    RTR_CHECK(cudaSetDevice(ctx->device), "native");
    float4 *d_mk = nullptr, *d_sk = nullptr;
    if (int e = tmp_alloc(ctx, &d_mk, km, "native")) return e;
    if (int e = tmp_alloc(ctx, &d_sk, ks, "native")) return e;
    if (km) RTR_CHECK(cudaMemcpyAsync(d_mk, host_model_kp_xyz1, (size_t)km * 16, cudaMemcpyHostToDevice, ctx->stream), "native");
    if (ks) RTR_CHECK(cudaMemcpyAsync(d_sk, host_scan_kp_xyz1, (size_t)ks * 16, cudaMemcpyHostToDevice, ctx->stream), "native");
    NatDesc dm, ds;
    if (int e = nat_describe(model, d_mk, km, p, true, dm)) return e;
    if (int e = nat_describe(scan, d_sk, ks, p, false, ds)) return e;
    float *score = nullptr, *transform = nullptr; int* best = nullptr;
    if (int e = nat_pairs(ctx, dm, ds, p, &score, &best, &transform)) return e;
    size_t np = (size_t)km * ks;
    if (np) {
        if (host_score) RTR_CHECK(cudaMemcpyAsync(host_score, score, np * 4, cudaMemcpyDeviceToHost, ctx->stream), "native");
        if (host_best_step) RTR_CHECK(cudaMemcpyAsync(host_best_step, best, np * 4, cudaMemcpyDeviceToHost, ctx->stream), "native");
        if (host_transform16) RTR_CHECK(cudaMemcpyAsync(host_transform16, transform, np * 64, cudaMemcpyDeviceToHost, ctx->stream), "native");
    }
    RTR_CHECK(cudaStreamSynchronize(ctx->stream), "native");
    nat_free(ctx, dm); nat_free(ctx, ds);
    dev_free(ctx, score); dev_free(ctx, best); dev_free(ctx, transform); dev_free(ctx, d_mk); dev_free(ctx, d_sk);
    return 0;
}

int rtr_native_register(rtr_cloud* model, rtr_cloud* scan, const rtr_native_params* p, rtr_pose_result* host_result) {
    if (!model || !scan || !p || !host_result || model->ctx != scan->ctx) return rtr_fail("native", "bad argument", RTR_ERR_INVALID);
    if (int e = nat_validate(p, true)) return e;
    rtr_context* ctx = model->ctx;
    TmpScope tmp_scope(ctx);
    RTR_CHECK(cudaSetDevice(ctx->device), "native");
    // ModelPoint::getKeypoint / ScanPoint::getKeypoint (model_point.h:127-136): Harris r = 0.05, thr = 0.01, NMS, refine
    rtr_cloud* cl[2] = {model, scan};
    int* d_idx[2]; float4* d_xyz[2]; int* d_cnt[2];
    int h_cnt[2] = {0, 0};
    for (int i = 0; i < 2; ++i) {
        if (int e = rtr_normals_dev(cl[i], 0.05f)) return e;
        if (int e = rtr_harris_dev(cl[i], 0.05f, 0.01f, 1, 1, &d_idx[i], &d_xyz[i], &d_cnt[i])) return e;
        RTR_CHECK(cudaMemcpyAsync(&h_cnt[i], d_cnt[i], sizeof(int), cudaMemcpyDeviceToHost, ctx->stream), "native");
    }
    RTR_CHECK(cudaStreamSynchronize(ctx->stream), "native");       // the keypoint counts size the launches below
    int km = h_cnt[0], ks = h_cnt[1];
    NatDesc dm, ds;
    if (int e = nat_describe(model, d_xyz[0], km, p, true, dm)) return e;
    if (int e = nat_describe(scan, d_xyz[1], ks, p, false, ds)) return e;
    float *score = nullptr, *transform = nullptr; int *best = nullptr, *pair_list = nullptr;
    if (int e = nat_pairs(ctx, dm, ds, p, &score, &best, &transform)) return e;
    rtr_pose_result* d_res = nullptr;
    if (int e = tmp_alloc(ctx, &d_res, 1, "native")) return e;
    if (int e = tmp_alloc(ctx, &pair_list, (size_t)km * ks, "native")) return e;
    // modelpoint.getArea(mcloud) / scanpoint.get_Area(cloud) (RealTimeRobot.cpp:41,47): the kept surfaces of both clouds
    NatSurface* d_surf[2] = {nullptr, nullptr};
    int n_surf[2] = {0, 0};
    if (p->use_plane_areas) {
        for (int i = 0; i < 2; ++i) {
            std::vector<rtr_surface> all(256);
            int np = 0;
            int rcp = rtr_plane_areas(cl[i], all.data(), 256, &np);
            if (rcp != 0 && rcp != RTR_ERR_CAPACITY) return rcp;
            std::vector<NatSurface> kept;
            for (int j = 0; j < std::min(np, 256); ++j) if (all[j].kept) {
                NatSurface sf; memcpy(sf.c, all[j].coefficients, sizeof(sf.c)); sf.area = all[j].area; sf.is_vertical = all[j].is_vertical; sf.pad = 0;
                kept.push_back(sf);
            }
            n_surf[i] = (int)kept.size();
            if (int e = tmp_alloc(ctx, &d_surf[i], kept.size(), "native")) return e;
            if (!kept.empty()) {
                RTR_CHECK(cudaMemcpyAsync(d_surf[i], kept.data(), kept.size() * sizeof(NatSurface), cudaMemcpyHostToDevice, ctx->stream), "native");
                RTR_CHECK(cudaStreamSynchronize(ctx->stream), "native");      // `kept` is a stack temporary
            }
        }
    }
    k_native_consensus<<<1, 1024, 0, ctx->stream>>>(dm.kps, km, ds.kps, ks, dm.number, ds.number, score, transform, d_surf[0], n_surf[0],
                                                     d_surf[1], n_surf[1], *p, pair_list, d_res);
    RTR_LAUNCH_CHECK(ctx, "native.consensus");
    RTR_CHECK(cudaMemcpyAsync(ctx->pinned, d_res, sizeof(rtr_pose_result), cudaMemcpyDeviceToHost, ctx->stream), "native");
    RTR_CHECK(cudaStreamSynchronize(ctx->stream), "native");
    memcpy(host_result, ctx->pinned, sizeof(rtr_pose_result));
    nat_free(ctx, dm); nat_free(ctx, ds);
    dev_free(ctx, score); dev_free(ctx, best); dev_free(ctx, transform); dev_free(ctx, pair_list); dev_free(ctx, d_res);
    for (int i = 0; i < 2; ++i) { dev_free(ctx, d_idx[i]); dev_free(ctx, d_xyz[i]); dev_free(ctx, d_cnt[i]); }
    return 0;
}

}  // extern "C"
