// Instance loss with linear assignment, on the device.
//
// Replaces LinAssignmentThingsLoss (loss/lin_assignment_things.py:13-89; called at pc_nerf/trainer.py:484-520), which builds the
// label x id cost matrix in a Python loop, moves it to the host (.cpu() per label), solves it with
// scipy.optimize.linear_sum_assignment and relabels in another Python loop -- one device-host round trip per label and per
// image in every training step.  Here one step is five launches and no synchronisation:
//   labels   one CTA per image: bitonic sort of the ground-truth ids in shared memory -> sorted unique "things" labels (first C-1,
//            :31) and, per ray, the rank of its label in that list
//   cost     warp per ray: sum of the predicted probabilities of classes 1.. per label (:33-34), counts, optional sum of x (:41-45)
//   assign   one CTA per image: cost = -(sum / (count + 1e-4)) (:34), optional id-range rejection (utils/outlier_rejection.py:8-52),
//            nan_to_num, shortest-augmenting-path assignment (rows = labels <= columns = ids), one thread per column
//   virtual  warp per ray: virtual label (:49-55), arg-max prediction, per-image "anything wrong?" flag (:84)
//   nll      loss[b, r] = -log(p[b, r, virtual] + 1e-27) on the valid rays of the flagged images (:85), and its backward
#include "common.cuh"
#include <float.h>

#define LOSS_MAX_IDS 256      // C - 1 <= 255 columns

// ---------------------------------------------------------------------------------------------
// labels: sort + unique + rank
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) inst_labels_kernel(const int64_t* __restrict__ gt, int64_t R, int Rpad, int m,
                                                           int* __restrict__ labels, int* __restrict__ n_labels, int* __restrict__ rank) {
    extern __shared__ int sh[];      // keys[Rpad] | uniq[m]
    int* keys = sh;
    int* uniq = sh + Rpad;
    __shared__ int n_s;
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    const int64_t* g = gt + (int64_t)b * R;
    for (int i = tid; i < Rpad; i += nt) {
        const int64_t v = i < R ? g[i] : 0;
        keys[i] = v > 0 ? (int)v : INT_MAX;
    }
    __syncthreads();
    for (int k = 2; k <= Rpad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < Rpad; i += nt) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const int a = keys[i], c = keys[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > c) == up) { keys[i] = c; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    if (tid == 0) {      // unique heads, in order (<= a few hundred distinct labels: serial is fine next to the sort)
        int n = 0;
        for (int i = 0; i < Rpad && keys[i] != INT_MAX; ++i)
            if (i == 0 || keys[i] != keys[i - 1]) {
                if (n < m) uniq[n] = keys[i];
                ++n;
            }
        n_s = n < m ? n : m;
        n_labels[b] = n_s;
    }
    __syncthreads();
    const int n = n_s;
    for (int i = tid; i < m; i += nt) labels[b * m + i] = i < n ? uniq[i] : 0;
    for (int i = tid; i < R; i += nt) {
        const int64_t v = g[i];
        int r = -1;                       // not a "thing"
        if (v > 0) {
            int lo = 0, hi = n;           // first index with uniq >= v
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (uniq[mid] < (int)v) lo = mid + 1; else hi = mid; }
            r = (lo < n && uniq[lo] == (int)v) ? lo : -2;      // -2: label beyond the first C-1 (keeps things_labels = 0, :48)
        }
        rank[(int64_t)b * R + i] = r;
    }
}

// ---------------------------------------------------------------------------------------------
// cost accumulation: warp per ray
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) inst_cost_kernel(const float* __restrict__ p, const int* __restrict__ rank, const float* __restrict__ points,
                                                        int64_t B, int64_t R, int C, float* __restrict__ csum, float* __restrict__ cnt,
                                                        float* __restrict__ xsum) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= B * R) return;
    const int r = rank[w];
    if (r < 0) return;
    const int64_t b = w / R;
    const int m = C - 1;
    const float* row = p + w * C + 1;
    float* dst = csum + ((int64_t)b * m + r) * m;
    for (int c = lane; c < m; c += 32) red_add_f32(dst + c, __ldg(row + c));
    if (lane == 0) {
        red_add_f32(cnt + b * m + r, 1.f);
        if (points) red_add_f32(xsum + b * m + r, __ldg(points + w * 3));
    }
}

// ---------------------------------------------------------------------------------------------
// assignment: shortest augmenting paths, one CTA per image, one thread per column (1-based columns as in the classic formulation)
// ---------------------------------------------------------------------------------------------
struct ArgMin { double v; int j; };
__device__ __forceinline__ ArgMin argmin_combine(ArgMin a, ArgMin b) { return (b.v < a.v || (b.v == a.v && b.j < a.j)) ? b : a; }

__global__ void __launch_bounds__(256) inst_assign_kernel(const float* __restrict__ csum, const float* __restrict__ cnt, const float* __restrict__ xsum,
                                                          const int* __restrict__ n_labels, int C, int outlier, float frame_min_length,
                                                          int max_num_inst_at_x, int id_margin, int* __restrict__ assign) {
    extern __shared__ float cost[];            // [n][m]
    __shared__ double u[LOSS_MAX_IDS + 1], v[LOSS_MAX_IDS + 1], minv[LOSS_MAX_IDS + 1];
    __shared__ int pj[LOSS_MAX_IDS + 1], way[LOSS_MAX_IDS + 1];
    __shared__ unsigned char used[LOSS_MAX_IDS + 1];
    __shared__ ArgMin red[8];
    __shared__ int j0_s, j1_s;
    __shared__ double delta_s;
    const int b = blockIdx.x, tid = threadIdx.x, m = C - 1, n = n_labels[b];
    for (int i = tid; i < n * m; i += blockDim.x) {
        const int l = i / m, c = i - l * m;
        float x = -(csum[((int64_t)b * m + l) * m + c] / (cnt[b * m + l] + 1e-4f));
        if (outlier) {      // ids outside the range available at the label's mean x position cost 10000 (outlier_rejection.py:8-52)
            const float cx = xsum[b * m + l] / cnt[b * m + l];
            const float slope = (float)(max_num_inst_at_x + id_margin) / frame_min_length;
            const float x_limit = (float)(m - id_margin) / slope;
            float xr = (-cx + 1.f) * 0.5f;
            xr = xr - floorf(xr / x_limit) * x_limit;                    // python % for a positive modulus
            long long lo = (long long)fminf(fmaxf(slope * xr, 0.f), (float)(m - 1));
            long long hi = lo + id_margin; if (hi > m - 1) hi = m - 1; if (hi < 0) hi = 0;
            if (!(lo <= c && c <= hi)) x = 10000.f;
        }
        if (isnan(x)) x = 0.f; else if (isinf(x)) x = x > 0.f ? FLT_MAX : -FLT_MAX;      // np.nan_to_num
        cost[i] = x;
    }
    for (int j = tid; j <= m; j += blockDim.x) { u[j] = 0.0; v[j] = 0.0; pj[j] = 0; way[j] = 0; }
    __syncthreads();
    for (int i = 1; i <= n; ++i) {
        if (tid == 0) { pj[0] = i; j0_s = 0; }
        for (int j = tid; j <= m; j += blockDim.x) { minv[j] = DBL_MAX; used[j] = 0; }
        __syncthreads();
        while (true) {
            const int j0 = j0_s;
            if (tid == 0) used[j0] = 1;
            __syncthreads();
            const int i0 = pj[j0];
            ArgMin best{DBL_MAX, INT_MAX};
            for (int j = 1 + tid; j <= m; j += blockDim.x) {
                if (!used[j]) {
                    const double cur = (double)cost[(i0 - 1) * m + (j - 1)] - u[i0] - v[j];
                    if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
                    if (minv[j] < best.v || (minv[j] == best.v && j < best.j)) { best.v = minv[j]; best.j = j; }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ArgMin other;
                other.v = __shfl_xor_sync(0xffffffffu, best.v, o);
                other.j = __shfl_xor_sync(0xffffffffu, best.j, o);
                best = argmin_combine(best, other);
            }
            if ((tid & 31) == 0) red[tid >> 5] = best;
            __syncthreads();
            if (tid == 0) {
                ArgMin t = red[0];
                for (int k = 1; k < (int)(blockDim.x >> 5); ++k) t = argmin_combine(t, red[k]);
                delta_s = t.v; j1_s = t.j;
            }
            __syncthreads();
            const double delta = delta_s;
            for (int j = tid; j <= m; j += blockDim.x) {
                if (used[j]) { u[pj[j]] += delta; v[j] -= delta; }      // distinct used columns carry distinct rows: no collision
                else minv[j] -= delta;
            }
            __syncthreads();
            if (tid == 0) j0_s = j1_s;
            __syncthreads();
            if (pj[j0_s] == 0) break;
        }
        if (tid == 0) {
            int j0 = j0_s;
            do { const int j1 = way[j0]; pj[j0] = pj[j1]; j0 = j1; } while (j0);
        }
        __syncthreads();
    }
    for (int j = 1 + tid; j <= m; j += blockDim.x)
        if (pj[j] > 0) assign[b * m + (pj[j] - 1)] = j - 1;
}

// ---------------------------------------------------------------------------------------------
// virtual labels + arg-max + per-image flag; loss and its backward
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) inst_virtual_kernel(const float* __restrict__ p, const int64_t* __restrict__ gt, const uint8_t* __restrict__ stuff,
                                                           const int* __restrict__ rank, const int* __restrict__ assign, int64_t B, int64_t R, int C,
                                                           int* __restrict__ virt, int* __restrict__ flag) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= B * R) return;
    const int64_t b = w / R;
    const bool valid = stuff[w] || gt[w] > 0;
    if (!valid) { if (lane == 0) virt[w] = -1; return; }
    const int r = rank[w];
    const int vl = r >= 0 ? assign[b * (C - 1) + r] + 1 : (r == -2 ? 1 : 0);
    const float* row = p + w * C;
    float best = -INFINITY; int bi = INT_MAX;
    for (int c = lane; c < C; c += 32) { const float x = __ldg(row + c); if (x > best) { best = x; bi = c; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) {
        virt[w] = vl;
        if (bi != vl) atomicOr(flag + b, 1);
    }
}
__global__ void inst_nll_kernel(const float* __restrict__ p, const int* __restrict__ virt, const int* __restrict__ flag, int64_t B, int64_t R, int C,
                                float* __restrict__ loss) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= B * R) return;
    const int vl = virt[w];
    loss[w] = (vl >= 0 && flag[w / R]) ? -logf(p[w * C + vl] + 1e-27f) : 0.f;
}
__global__ void inst_nll_bwd_kernel(const float* __restrict__ p, const int* __restrict__ virt, const int* __restrict__ flag, const float* __restrict__ g,
                                    int64_t B, int64_t R, int C, float* __restrict__ gp) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= B * R) return;
    const int vl = virt[w];
    if (vl >= 0 && flag[w / R]) gp[w * C + vl] = -g[w] / (p[w * C + vl] + 1e-27f);
}

// ---------------------------------------------------------------------------------------------
// photometric + panoptic NLL loss of a training step in one launch (and one for its gradients)
// ---------------------------------------------------------------------------------------------
// loss = w_rgb * mean |rgb - t_rgb| + w_sem * mean -log(sem[n, t_sem[n]] + eps) + w_inst * mean -log(inst[n, t_inst[n]] + eps)
// (pc_nerf/trainer.py:442-480: L1 rgb, NLL of log(p + 1e-27) for the composited semantic / instance probabilities).  As torch ops
// this is ~20 launches of a few microseconds each between the forward and the backward of a 1.2 ms step.
__global__ void __launch_bounds__(256) panoptic_loss_fwd_kernel(const float* __restrict__ rgb, const float* __restrict__ sem, const float* __restrict__ inst,
                                                                const float* __restrict__ t_rgb, const int64_t* __restrict__ t_sem,
                                                                const int64_t* __restrict__ t_inst, int64_t N, int Cs, int Ci, float w_rgb, float w_sem,
                                                                float w_inst, float eps, float* __restrict__ partials, unsigned* __restrict__ ticket,
                                                                float* __restrict__ loss) {
    float acc = 0.f;
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        if (rgb) acc += w_rgb * (fabsf(rgb[3 * n] - t_rgb[3 * n]) + fabsf(rgb[3 * n + 1] - t_rgb[3 * n + 1]) + fabsf(rgb[3 * n + 2] - t_rgb[3 * n + 2])) * (1.f / 3.f);
        if (sem) acc -= w_sem * logf(sem[n * Cs + t_sem[n]] + eps);
        if (inst) acc -= w_inst * logf(inst[n * Ci + t_inst[n]] + eps);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ float ws[8];
    __shared__ bool last;
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int k = 0; k < 8; ++k) t += ws[k];
        partials[blockIdx.x] = t;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {      // deterministic: the last block adds the partial sums in block order
        __threadfence();
        float t = 0.f;
        for (unsigned k = 0; k < gridDim.x; ++k) t += *reinterpret_cast<volatile float*>(partials + k);
        *loss = t / (float)N;
        *ticket = 0u;
    }
}
// warp per ray: the wide rows are written coalesced (zeros + the one non-zero of the NLL)
__global__ void __launch_bounds__(256) panoptic_loss_bwd_kernel(const float* __restrict__ rgb, const float* __restrict__ sem, const float* __restrict__ inst,
                                                                const float* __restrict__ t_rgb, const int64_t* __restrict__ t_sem,
                                                                const int64_t* __restrict__ t_inst, int64_t N, int Cs, int Ci, float w_rgb, float w_sem,
                                                                float w_inst, float eps, const float* __restrict__ g_loss, float* __restrict__ g_rgb,
                                                                float* __restrict__ g_sem, float* __restrict__ g_inst) {
    const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    const float g = __ldg(g_loss) / (float)N;
    if (g_rgb && lane < 3) {
        const float d = rgb[3 * n + lane] - t_rgb[3 * n + lane];
        g_rgb[3 * n + lane] = g * w_rgb * (1.f / 3.f) * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
    }
    if (g_sem) {
        const int t = (int)t_sem[n];
        for (int c = lane; c < Cs; c += 32) g_sem[n * Cs + c] = (c == t) ? -g * w_sem / (sem[n * Cs + t] + eps) : 0.f;
    }
    if (g_inst) {
        const int t = (int)t_inst[n];
        for (int c = lane; c < Ci; c += 32) g_inst[n * Ci + c] = (c == t) ? -g * w_inst / (inst[n * Ci + t] + eps) : 0.f;
    }
}

extern "C" {

// rgb f32[N,3] / sem f32[N,Cs] / inst f32[N,Ci] (each nullable with its target); partials f32[>= 148*4], ticket u32[1] zero on entry
// (left zero); loss f32[1].  Backward: g_loss f32[1] on the device; gradient tensors fully written (nullable per channel).
int pag_panoptic_loss_fwd(const float* rgb, const float* sem, const float* inst, const float* t_rgb, const int64_t* t_sem, const int64_t* t_inst,
                          int64_t N, int Cs, int Ci, float w_rgb, float w_sem, float w_inst, float eps, float* partials, uint32_t* ticket,
                          float* loss, void* stream) {
    if (N <= 0) return PAG_ERR_ARG;
    int grid = (int)((N + 255) / 256);
    if (grid > 148 * 4) grid = 148 * 4;
    panoptic_loss_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(rgb, sem, inst, t_rgb, t_sem, t_inst, N, Cs, Ci, w_rgb, w_sem, w_inst, eps,
                                                                    partials, ticket, loss);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}
int pag_panoptic_loss_bwd(const float* rgb, const float* sem, const float* inst, const float* t_rgb, const int64_t* t_sem, const int64_t* t_inst,
                          int64_t N, int Cs, int Ci, float w_rgb, float w_sem, float w_inst, float eps, const float* g_loss, float* g_rgb,
                          float* g_sem, float* g_inst, void* stream) {
    if (N <= 0) return PAG_ERR_ARG;
    panoptic_loss_bwd_kernel<<<pag_grid(N * 32, 256), 256, 0, (cudaStream_t)stream>>>(rgb, sem, inst, t_rgb, t_sem, t_inst, N, Cs, Ci, w_rgb, w_sem,
                                                                                     w_inst, eps, g_loss, g_rgb, g_sem, g_inst);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

// p f32[B, R, C] probabilities, gt i64[B, R], stuff u8[B, R], points f32[B, R, 3] (nullable: no outlier rejection).
// Workspaces (caller allocates; csum / cnt / xsum / flag zeroed): labels i32[B, C-1], n_labels i32[B], rank i32[B, R],
// csum f32[B, C-1, C-1], cnt f32[B, C-1], xsum f32[B, C-1], assign i32[B, C-1], virt i32[B, R], flag i32[B].  loss f32[B, R].
int pag_inst_assignment_loss_fwd(const float* p, const int64_t* gt, const uint8_t* stuff, const float* points, int64_t B, int64_t R, int C,
                                 float frame_min_length, int max_num_inst_at_x, int id_margin, int32_t* labels, int32_t* n_labels, int32_t* rank,
                                 float* csum, float* cnt, float* xsum, int32_t* assign, int32_t* virt, int32_t* flag, float* loss, void* stream) {
    if (C < 2 || C - 1 > LOSS_MAX_IDS - 1 || B < 0 || R < 0) return PAG_ERR_UNSUPPORTED;
    if (B == 0 || R == 0) return PAG_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int m = C - 1;
    int Rpad = 1;
    while (Rpad < R) Rpad <<= 1;
    const size_t sm_labels = (size_t)(Rpad + m) * sizeof(int);
    if (sm_labels > 200 * 1024) return PAG_ERR_UNSUPPORTED;      // rays per image <= ~50 k
    cudaError_t e = cudaFuncSetAttribute(inst_labels_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_labels);
    if (e != cudaSuccess) return (int)e;
    inst_labels_kernel<<<(unsigned)B, 1024, sm_labels, st>>>(gt, R, Rpad, m, labels, n_labels, rank);
    PAG_LAUNCH_CHECK();
    inst_cost_kernel<<<pag_grid(B * R * 32, 256), 256, 0, st>>>(p, rank, points, B, R, C, csum, cnt, xsum);
    PAG_LAUNCH_CHECK();
    const size_t sm_cost = (size_t)m * m * sizeof(float);
    e = cudaFuncSetAttribute(inst_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_cost);
    if (e != cudaSuccess) return (int)e;
    inst_assign_kernel<<<(unsigned)B, 256, sm_cost, st>>>(csum, cnt, xsum, n_labels, C, points != nullptr, frame_min_length, max_num_inst_at_x,
                                                         id_margin, assign);
    PAG_LAUNCH_CHECK();
    inst_virtual_kernel<<<pag_grid(B * R * 32, 256), 256, 0, st>>>(p, gt, stuff, rank, assign, B, R, C, virt, flag);
    PAG_LAUNCH_CHECK();
    inst_nll_kernel<<<pag_grid(B * R, 256), 256, 0, st>>>(p, virt, flag, B, R, C, loss);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}
// gp f32[B, R, C] zeroed by the caller; one non-zero per valid ray of a flagged image
int pag_inst_assignment_loss_bwd(const float* p, const int32_t* virt, const int32_t* flag, const float* g_loss, int64_t B, int64_t R, int C,
                                 float* gp, void* stream) {
    if (B == 0 || R == 0) return PAG_OK;
    inst_nll_bwd_kernel<<<pag_grid(B * R, 256), 256, 0, (cudaStream_t)stream>>>(p, virt, flag, g_loss, B, R, C, gp);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

}  // extern "C"
