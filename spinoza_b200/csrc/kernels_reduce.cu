// kernels_reduce.cu -- measurement / probability reductions, collapse, state initialisation, sampling.
//
// Replaces measurement.rs:12-92 (measure_qubit), core.rs:198-264 (expectation values) and
// core.rs:65-129 (sampling).  All reductions are deterministic two-stage sums: per-thread serial
// accumulation over a grid-stride loop, warp shuffle, one partial per CTA, then a single-CTA final sum
// (the reference's rayon reduction order is nondeterministic; parity is to 1e-12, not bitwise).
// Read-only streaming passes: 16 * 2^n bytes for norm2 / <X>,<Y>,<Z>; 8 * 2^n for prob0 (the s0 half).
#include <algorithm>
#include <vector>

#include "gate_math.cuh"
#include "kernels_zall.cuh"
#include "kernels_xyall.cuh"

namespace spz {

constexpr int kRedThreads = 256;
constexpr int kRedBlocks = 148 * 8; // grid sized in multiples of the SM count
constexpr int kZGrid = 148 * 2;     // the all-qubit <Z> pass: two CTAs per SM (its accumulators take ~100 registers per thread)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double block_sum(double v) {
    __shared__ double sh[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        v = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
        v = warp_sum(v);
    }
    __syncthreads();
    return v; // valid in warp 0
}

int ensure_scratch(spz_state *st) {
    if (!st->scratch.partials) {
        // [0, kRedBlocks + 8): one-target reductions.  Behind it: the all-qubit <Z> pass, (kZMaxBits + 1) values per CTA + its result
        SPZ_CUDA(cudaMalloc(&st->scratch.partials, sizeof(double) * (kRedBlocks + 8 + (size_t)(kZGrid + 1) * (kZMaxBits + 1))));
        st->scratch.n_partials = kRedBlocks + 8 + (size_t)(kZGrid + 1) * (kZMaxBits + 1);
    }
    if (!st->scratch.h_result) SPZ_CUDA(cudaMallocHost(&st->scratch.h_result, sizeof(double) * 64));
    return SPZ_OK;
}

// mode 0: sum |amp|^2 over amps with target bit 0 (measurement.rs:16-29)
// mode 1: sum |amp|^2 over all amps
// mode 2/3/4: Re<psi|O psi> for O = X/Y/Z on `target` (core.rs:239-258 without the state clone):
//   X: v[s0]=s1, v[s1]=s0            -> sum 2(ac + bd)
//   Y: v[s0]=(d,-c), v[s1]=(-b,a)    -> sum 2(ad - bc)
//   Z: v[s1]=-s1                     -> sum (a^2+b^2) - (c^2+d^2)
template <int MODE>
__global__ void __launch_bounds__(kRedThreads) k_reduce(const double *__restrict__ re, const double *__restrict__ im,
                                                       long long npairs, int target, double *__restrict__ partials) {
    double acc = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    if constexpr (MODE == 1) {
        // npairs = number of double2 vectors (len/2) or len when len < 2
        const long long nvec = npairs;
        const double2 *r2 = reinterpret_cast<const double2 *>(re);
        const double2 *m2 = reinterpret_cast<const double2 *>(im);
        for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
            const double2 r = r2[v], m = m2[v];
            acc += r.x * r.x + m.x * m.x;
            acc += r.y * r.y + m.y * m.y;
        }
    } else {
        for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += stride) {
            const unsigned long long s0 = insert_zero((unsigned long long)p, target);
            const double a = re[s0], b = im[s0];
            if constexpr (MODE == 0) {
                acc += a * a + b * b;
            } else {
                const unsigned long long s1 = s0 | (1ull << target);
                const double c = re[s1], d = im[s1];
                if constexpr (MODE == 2) acc += 2.0 * (a * c + b * d);
                if constexpr (MODE == 3) acc += 2.0 * (a * d - b * c);
                if constexpr (MODE == 4) acc += (a * a + b * b) - (c * c + d * d);
            }
        }
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(kRedThreads) k_final_sum(const double *__restrict__ partials, int n, double *out) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partials[i];
    acc = block_sum(acc);
    if (threadIdx.x == 0) *out = acc;
}

int reduce_scalar(spz_state *st, int mode, int target, double *out) {
    SPZ_TRY(join_pending(st));
    SPZ_TRY(ensure_scratch(st));
    const long long len = st->len;
    long long work = mode == 1 ? (len >= 2 ? len / 2 : 0) : len / 2;
    if (mode != 1 && (target < 0 || target >= st->n)) { set_error("target %d out of range", target); return SPZ_ERR_INVALID_ARG; }
    int grid = (int)std::min<long long>((work + kRedThreads - 1) / kRedThreads, kRedBlocks);
    if (grid < 1) grid = 1;
    double *part = st->scratch.partials;
    switch (mode) {
    case 0: k_reduce<0><<<grid, kRedThreads, 0, st->stream>>>(st->re, st->im, work, target, part); break;
    case 1: k_reduce<1><<<grid, kRedThreads, 0, st->stream>>>(st->re, st->im, work, 0, part); break;
    case 2: k_reduce<2><<<grid, kRedThreads, 0, st->stream>>>(st->re, st->im, work, target, part); break;
    case 3: k_reduce<3><<<grid, kRedThreads, 0, st->stream>>>(st->re, st->im, work, target, part); break;
    case 4: k_reduce<4><<<grid, kRedThreads, 0, st->stream>>>(st->re, st->im, work, target, part); break;
    default: return SPZ_ERR_INVALID_ARG;
    }
    k_final_sum<<<1, kRedThreads, 0, st->stream>>>(part, grid, part + kRedBlocks);
    count_launch(2);
    SPZ_CUDA(cudaGetLastError());
    SPZ_CUDA(cudaMemcpyAsync(st->scratch.h_result, part + kRedBlocks, sizeof(double), cudaMemcpyDeviceToHost, st->stream));
    SPZ_CUDA(cudaStreamSynchronize(st->stream));
    *out = st->scratch.h_result[0];
    return SPZ_OK;
}

// ---- <Z_t> for every qubit at once (kernels_zall.cuh) -------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(kRedThreads, 2) k_z_all(const double *__restrict__ re, const double *__restrict__ im, long long nvec,
                                                          int n, double *__restrict__ partials) {
    double total = 0.0, s1[NB];
#pragma unroll
    for (int t = 0; t < NB; ++t) s1[t] = 0.0;
    z_all_accumulate<NB>(re, im, nvec, n, (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x, total, s1);
    double *mine = partials + (size_t)blockIdx.x * (kZMaxBits + 1);
    total = block_sum(total);
    if (threadIdx.x == 0) mine[0] = total;
#pragma unroll
    for (int t = 0; t < NB; ++t) {
        if (t < n) { // (uniform over the CTA)
            const double v = block_sum(s1[t]);
            if (threadIdx.x == 0) mine[1 + t] = v;
        }
    }
}
// block j: sum of value j over the CTAs of k_z_all
__global__ void __launch_bounds__(kRedThreads) k_z_final(const double *__restrict__ partials, int n_cta, double *__restrict__ out) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < n_cta; i += blockDim.x) acc += partials[(size_t)i * (kZMaxBits + 1) + blockIdx.x];
    acc = block_sum(acc);
    if (threadIdx.x == 0) out[blockIdx.x] = acc;
}

// out[0] = sum |amp|^2 over the handle's amplitudes, out[1 + t] = the part of it at indices with bit t set, t < st->n.
// One read pass (needs st->n >= 2: a vector is four amplitudes).
int reduce_z_all(spz_state *st, double *out) {
    SPZ_TRY(join_pending(st));
    SPZ_TRY(ensure_scratch(st));
    const int n = st->n;
    if (n < 2 || n > kZMaxBits) { set_error("all-qubit <Z> pass needs 2..%d qubits", kZMaxBits); return SPZ_ERR_INVALID_ARG; }
    const long long nvec = st->len / 4;
    const int grid = (int)std::max<long long>(1, std::min<long long>((nvec + kRedThreads - 1) / kRedThreads, kZGrid));
    double *part = st->scratch.partials + kRedBlocks + 8;
    double *res = part + (size_t)kZGrid * (kZMaxBits + 1);
    if (n <= 16) k_z_all<16><<<grid, kRedThreads, 0, st->stream>>>(st->re, st->im, nvec, n, part);
    else if (n <= 24) k_z_all<24><<<grid, kRedThreads, 0, st->stream>>>(st->re, st->im, nvec, n, part);
    else if (n <= 32) k_z_all<32><<<grid, kRedThreads, 0, st->stream>>>(st->re, st->im, nvec, n, part);
    else k_z_all<kZMaxBits><<<grid, kRedThreads, 0, st->stream>>>(st->re, st->im, nvec, n, part);
    k_z_final<<<n + 1, kRedThreads, 0, st->stream>>>(part, grid, res);
    count_launch(2);
    SPZ_CUDA(cudaGetLastError());
    SPZ_CUDA(cudaMemcpyAsync(st->scratch.h_result, res, sizeof(double) * (n + 1), cudaMemcpyDeviceToHost, st->stream));
    SPZ_CUDA(cudaStreamSynchronize(st->stream));
    for (int i = 0; i <= n; ++i) out[i] = st->scratch.h_result[i];
    return SPZ_OK;
}

// ---- <X_t> / <Y_t> for up to twelve qubits at once (kernels_xyall.cuh) ------------------------------------------------
__global__ void __launch_bounds__(kXYThreads, 2) k_xy_all(const XYArgs a) {
    extern __shared__ __align__(16) unsigned char xy_smem[];
    double *sre = reinterpret_cast<double *>(xy_smem);
    double *sim = sre + (1u << (a.L + a.H));
    double acc[kXYBits];
#pragma unroll
    for (int b = 0; b < kXYBits; ++b) acc[b] = 0.0;
    __shared__ unsigned long long hoff[64];
    xy_prepare(a, hoff);
    for (long long t = blockIdx.x; t < a.n_tiles; t += gridDim.x) xy_tile_accumulate(a, t, hoff, sre, sim, acc);
#pragma unroll
    for (int b = 0; b < kXYBits; ++b) {
        if ((a.tmask >> b) & 1u) { // (uniform over the CTA)
            const double v = block_sum(acc[b]);
            if (threadIdx.x == 0) a.partials[(size_t)blockIdx.x * kXYBits + b] = v;
        }
    }
}
__global__ void __launch_bounds__(kRedThreads) k_xy_final(const double *__restrict__ partials, int n_cta, unsigned tmask, double *__restrict__ out) {
    double acc = 0.0;
    if ((tmask >> blockIdx.x) & 1u)
        for (int i = threadIdx.x; i < n_cta; i += blockDim.x) acc += partials[(size_t)i * kXYBits + blockIdx.x];
    acc = block_sum(acc);
    if (threadIdx.x == 0) out[blockIdx.x] = 2.0 * acc;
}

// <X> (obs 0) or <Y> (obs 1) of k targets, grouped into tiles: the targets below bit 12 share one pass (a tile of 2^12
// contiguous amplitudes), the higher ones go six at a time (64 contiguous amplitudes x 2^6 combinations of the six qubits).
int reduce_xy_multi(spz_state *st, int obs, const int32_t *targets, int k, double *out) {
    SPZ_TRY(join_pending(st));
    SPZ_TRY(ensure_scratch(st));
    const int n = st->n;
    for (int i = 0; i < k; ++i)
        if (targets[i] < 0 || targets[i] >= n) { set_error("target %d out of range", targets[i]); return SPZ_ERR_INVALID_ARG; }
    static bool attr_set[64] = {false}; // 64 KB of dynamic shared memory needs the opt-in, once per device
    if (st->device >= 0 && st->device < 64 && !attr_set[st->device]) {
        SPZ_CUDA(cudaFuncSetAttribute(k_xy_all, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2u * kXYTile * sizeof(double))));
        attr_set[st->device] = true;
    }
    uint64_t want = 0;
    for (int i = 0; i < k; ++i) want |= 1ull << targets[i];
    double value[64];
    double *part = st->scratch.partials + kRedBlocks + 8; // (the all-qubit <Z> pass uses the same region; never at the same time)
    double *res = part + (size_t)kZGrid * (kZMaxBits + 1);
    static_assert(kXYBits <= kZMaxBits + 1, "the <X>/<Y> partials fit the region sized for the <Z> pass");
    auto run = [&](const XYArgs &proto, const int *tile_qubit) -> int { // tile_qubit[b]: the qubit tile bit b stands for
        XYArgs a = proto;
        a.re = st->re; a.im = st->im; a.obs = obs; a.partials = part;
        const int tb = a.L + a.H;
        a.n_tiles = (long long)st->len >> tb;
        constexpr int kXYGrid = 148 * 3; // three CTAs per SM: 64 KB of staging each
        static_assert((size_t)kXYGrid * kXYBits <= (size_t)kZGrid * (kZMaxBits + 1), "partials region");
        const int grid = (int)std::max<long long>(1, std::min<long long>(a.n_tiles, kXYGrid));
        const size_t smem = 2u * sizeof(double) << tb;
        k_xy_all<<<grid, kXYThreads, smem, st->stream>>>(a);
        k_xy_final<<<kXYBits, kRedThreads, 0, st->stream>>>(part, grid, a.tmask, res);
        count_launch(2);
        SPZ_CUDA(cudaGetLastError());
        SPZ_CUDA(cudaMemcpyAsync(st->scratch.h_result, res, sizeof(double) * kXYBits, cudaMemcpyDeviceToHost, st->stream));
        SPZ_CUDA(cudaStreamSynchronize(st->stream));
        for (int b = 0; b < tb; ++b)
            if ((a.tmask >> b) & 1u) value[tile_qubit[b]] = st->scratch.h_result[b];
        return SPZ_OK;
    };
    const int low_bits = std::min(n, kXYBits);
    if (want & ((1ull << low_bits) - 1ull)) { // the targets inside one tile of contiguous amplitudes
        XYArgs a{};
        a.L = low_bits; a.H = 0;
        a.tmask = (unsigned)(want & ((1ull << low_bits) - 1ull));
        int tq[kXYBits];
        for (int b = 0; b < kXYBits; ++b) tq[b] = b;
        SPZ_TRY(run(a, tq));
    }
    std::vector<int> high;
    for (int q = low_bits; q < n; ++q) if ((want >> q) & 1ull) high.push_back(q);
    for (size_t i = 0; i < high.size(); i += 6) { // six high qubits per pass, on top of 64 contiguous amplitudes
        XYArgs a{};
        a.L = 6;
        a.H = (int)std::min<size_t>(6, high.size() - i);
        int tq[kXYBits] = {0};
        for (int h = 0; h < a.H; ++h) { a.high[h] = (unsigned char)high[i + h]; tq[a.L + h] = high[i + h]; a.tmask |= 1u << (a.L + h); }
        SPZ_TRY(run(a, tq));
    }
    for (int i = 0; i < k; ++i) out[i] = value[targets[i]];
    return SPZ_OK;
}

// ---- collapse (measurement.rs:39-90) --------------------------------------------------------------------
// outcome 0: s0 *= k, s1 = 0.   outcome 1: s1 *= k, s0 = 0; with reset the following X (measurement.rs:87-89)
// is folded into the same pass: s0 = s1 * k, s1 = 0 (X is an exact swap, so this is bit-identical).
__global__ void __launch_bounds__(256) k_collapse(double *__restrict__ re, double *__restrict__ im, long long npairs,
                                                  int target, int outcome, int reset, double k) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += stride) {
        const unsigned long long s0 = insert_zero((unsigned long long)p, target);
        const unsigned long long s1 = s0 | (1ull << target);
        if (outcome == 0) {
            re[s0] = __dmul_rn(re[s0], k);
            im[s0] = __dmul_rn(im[s0], k);
            re[s1] = 0.0;
            im[s1] = 0.0;
        } else {
            const double c = __dmul_rn(re[s1], k), d = __dmul_rn(im[s1], k);
            if (reset) {
                re[s0] = c; im[s0] = d; re[s1] = 0.0; im[s1] = 0.0;
            } else {
                re[s1] = c; im[s1] = d; re[s0] = 0.0; im[s0] = 0.0;
            }
        }
    }
}

int launch_collapse(spz_state *st, int target, int outcome, int reset, double scale) {
    SPZ_TRY(join_pending(st));
    const long long npairs = st->len / 2;
    const int grid = (int)std::max<long long>(1, std::min<long long>((npairs + 255) / 256, 148 * 16));
    k_collapse<<<grid, 256, 0, st->stream>>>(st->re, st->im, npairs, target, outcome, reset, scale);
    count_launch();
    SPZ_CUDA(cudaGetLastError());
    return SPZ_OK;
}

// ---- measurement inside a subspace ------------------------------------------------------------------------------
// A run of measurements with reset (QuantumCircuit::execute, circuit.rs:559-566: measure_qubit(state, target, true, None))
// leaves every measured qubit in |0>: all amplitudes whose index has one of those bits set are exactly 0.  The measurements
// that follow in the same run only have to visit the indices where those bits are 0 -- half as many after every measurement:
// measuring all n qubits reads ~4 * 2^n amplitudes in total instead of n * 2^n.  (pos: ascending bit positions to keep 0.)
struct SubArgs {
    int nins;
    unsigned char pos[kZMaxBits];
};
__device__ __forceinline__ unsigned long long sub_index(unsigned long long p, const SubArgs &a) {
    for (int k = 0; k < a.nins; ++k) p = insert_zero(p, a.pos[k]);
    return p;
}
// prob0 (measurement.rs:16-29) over the subspace: pos holds the known-zero qubits AND the target
__global__ void __launch_bounds__(kRedThreads) k_prob0_sub(const double *__restrict__ re, const double *__restrict__ im, long long count,
                                                           const SubArgs a, double *__restrict__ partials) {
    double acc = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < count; p += stride) {
        const unsigned long long s0 = sub_index((unsigned long long)p, a);
        const double x = re[s0], y = im[s0];
        acc += x * x + y * y;
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}
// the collapse of k_collapse over the same subspace (everything outside it is 0 and stays 0)
__global__ void __launch_bounds__(256) k_collapse_sub(double *__restrict__ re, double *__restrict__ im, long long count, const SubArgs a,
                                                      int target, int outcome, int reset, double k) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < count; p += stride) {
        const unsigned long long s0 = sub_index((unsigned long long)p, a);
        const unsigned long long s1 = s0 | (1ull << target);
        if (outcome == 0) {
            re[s0] = __dmul_rn(re[s0], k);
            im[s0] = __dmul_rn(im[s0], k);
            re[s1] = 0.0;
            im[s1] = 0.0;
        } else {
            const double c = __dmul_rn(re[s1], k), d = __dmul_rn(im[s1], k);
            if (reset) {
                re[s0] = c; im[s0] = d; re[s1] = 0.0; im[s1] = 0.0;
            } else {
                re[s1] = c; im[s1] = d; re[s0] = 0.0; im[s0] = 0.0;
            }
        }
    }
}
static bool sub_args(const spz_state *st, int target, uint64_t zero_mask, SubArgs *a) {
    if (target < 0 || target >= st->n || st->n > kZMaxBits) return false;
    int k = 0;
    for (int q = 0; q < st->n; ++q)
        if (q == target || ((zero_mask >> q) & 1ull)) a->pos[k++] = (unsigned char)q;
    a->nins = k;
    return true;
}
// sum |amp|^2 over the indices with bit `target` and every bit of zero_mask clear
int reduce_prob0_sub(spz_state *st, int target, uint64_t zero_mask, double *out) {
    SPZ_TRY(join_pending(st));
    SPZ_TRY(ensure_scratch(st));
    SubArgs a{};
    if (!sub_args(st, target, zero_mask, &a)) { set_error("target %d out of range", target); return SPZ_ERR_INVALID_ARG; }
    const long long count = (long long)st->len >> a.nins;
    const int grid = (int)std::max<long long>(1, std::min<long long>((count + kRedThreads - 1) / kRedThreads, kRedBlocks));
    double *part = st->scratch.partials;
    k_prob0_sub<<<grid, kRedThreads, 0, st->stream>>>(st->re, st->im, count, a, part);
    k_final_sum<<<1, kRedThreads, 0, st->stream>>>(part, grid, part + kRedBlocks);
    count_launch(2);
    SPZ_CUDA(cudaGetLastError());
    SPZ_CUDA(cudaMemcpyAsync(st->scratch.h_result, part + kRedBlocks, sizeof(double), cudaMemcpyDeviceToHost, st->stream));
    SPZ_CUDA(cudaStreamSynchronize(st->stream));
    *out = st->scratch.h_result[0];
    return SPZ_OK;
}
int launch_collapse_sub(spz_state *st, int target, int outcome, int reset, double scale, uint64_t zero_mask) {
    SPZ_TRY(join_pending(st));
    SubArgs a{};
    if (!sub_args(st, target, zero_mask, &a)) { set_error("target %d out of range", target); return SPZ_ERR_INVALID_ARG; }
    const long long count = (long long)st->len >> a.nins;
    const int grid = (int)std::max<long long>(1, std::min<long long>((count + 255) / 256, 148 * 16));
    k_collapse_sub<<<grid, 256, 0, st->stream>>>(st->re, st->im, count, a, target, outcome, reset, scale);
    count_launch();
    SPZ_CUDA(cudaGetLastError());
    return SPZ_OK;
}

// ---- initialisation -----------------------------------------------------------------------------------
__global__ void k_set_one(double *re, unsigned long long index) { re[index] = 1.0; }

int launch_fill_basis(spz_state *st, uint64_t index) {
    SPZ_TRY(join_pending(st));
    if (index >= (uint64_t)st->len) { set_error("basis index out of range"); return SPZ_ERR_INVALID_ARG; }
    SPZ_CUDA(cudaMemsetAsync(st->re, 0, sizeof(double) * (size_t)st->len, st->stream));
    SPZ_CUDA(cudaMemsetAsync(st->im, 0, sizeof(double) * (size_t)st->len, st->stream));
    k_set_one<<<1, 1, 0, st->stream>>>(st->re, index);
    count_launch();
    SPZ_CUDA(cudaGetLastError());
    return SPZ_OK;
}

// splitmix64 as a counter-based generator: the k-th output of the sequential generator seeded with
// `seed` is mix(seed + (k+1) * GOLDEN).  Same stream as oracle/spinoza_oracle.c:orc_gen_random_state.
__host__ __device__ __forceinline__ uint64_t splitmix_at(uint64_t seed, uint64_t k) {
    uint64_t z = seed + (k + 1ull) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ double u01_of(uint64_t r) { return (double)(r >> 11) * (1.0 / 9007199254740992.0); }

// gen_random_state utils.rs:168-201: p_i ~ U(0,1) normalised by their sum, phase ~ U(0, 2 pi).
// `offset` is the global index of this shard's first amplitude (0 on a single GPU).
__global__ void __launch_bounds__(256) k_rand_probs(double *__restrict__ re, long long len, long long offset, uint64_t seed,
                                                    double *__restrict__ partials) {
    double acc = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        const double u = u01_of(splitmix_at(seed, (uint64_t)(offset + i)));
        re[i] = u;
        acc += u;
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}
__global__ void __launch_bounds__(256) k_rand_finish(double *__restrict__ re, double *__restrict__ im, long long len,
                                                     long long offset, long long total_len, uint64_t seed,
                                                     const double *__restrict__ total) {
    const double recip = 1.0 / *total;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        const double p = re[i] * recip;
        const double ang = u01_of(splitmix_at(seed, (uint64_t)(total_len + offset + i))) * (2.0 * 3.14159265358979323846);
        const double ps = sqrt(p);
        double sn, cs;
        sincos(ang, &sn, &cs);
        re[i] = ps * cs;
        im[i] = ps * sn;
    }
}

int launch_rand_probs(spz_state *st, uint64_t seed, long long index_offset, double **d_local_total) {
    SPZ_TRY(join_pending(st));
    SPZ_TRY(ensure_scratch(st));
    const int grid = (int)std::max<long long>(1, std::min<long long>((st->len + 255) / 256, kRedBlocks));
    double *part = st->scratch.partials;
    k_rand_probs<<<grid, 256, 0, st->stream>>>(st->re, st->len, index_offset, seed, part);
    k_final_sum<<<1, kRedThreads, 0, st->stream>>>(part, grid, part + kRedBlocks);
    count_launch(2);
    SPZ_CUDA(cudaGetLastError());
    *d_local_total = part + kRedBlocks;
    return SPZ_OK;
}

int launch_rand_finish(spz_state *st, uint64_t seed, long long index_offset, long long total_len, const double *d_total) {
    const int grid = (int)std::max<long long>(1, std::min<long long>((st->len + 255) / 256, kRedBlocks));
    k_rand_finish<<<grid, 256, 0, st->stream>>>(st->re, st->im, st->len, index_offset, total_len, seed, d_total);
    count_launch();
    SPZ_CUDA(cudaGetLastError());
    return SPZ_OK;
}

int launch_init_random(spz_state *st, uint64_t seed) {
    double *d_total = nullptr;
    SPZ_TRY(launch_rand_probs(st, seed, 0, &d_total));
    return launch_rand_finish(st, seed, 0, st->len, d_total);
}

__global__ void __launch_bounds__(256) k_scale_all(double *__restrict__ re, double *__restrict__ im, long long len, double k) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        re[i] = __dmul_rn(re[i], k);
        im[i] = __dmul_rn(im[i], k);
    }
}

int launch_scale(spz_state *st, double scale) {
    SPZ_TRY(join_pending(st));
    const int grid = (int)std::max<long long>(1, std::min<long long>((st->len + 255) / 256, 148 * 16));
    k_scale_all<<<grid, 256, 0, st->stream>>>(st->re, st->im, st->len, scale);
    count_launch();
    SPZ_CUDA(cudaGetLastError());
    return SPZ_OK;
}

// ---- sampling ---------------------------------------------------------------------------------------------
// Exact inverse-CDF sampling (replaces the O(num_tests * k) reservoir loop of core.rs:81-112), two streaming read passes
// over the state however many shots are asked for:
//   pass 1   level 1: |amp|^2 summed per block of B = 4096 amplitudes (k_block_prob); its inclusive prefix per group of
//            4096 blocks (k_scan_groups) and the prefix over the <= 512 group totals (k_scan_single): a two-level CDF, on
//            the device;
//   locate   one thread per shot: x = u * total, two binary searches -> the block that holds the shot and the residual
//            inside it; a counting sort groups the shots by block (histogram, exclusive scan, scatter);
//   pass 2   one CTA per block that holds shots: the block's 4096 probabilities are scanned once in shared memory and every
//            shot of the bucket is a 12-step binary search there (k_sample_blocks).  Blocks without shots are not read.
// Every sum has a fixed order, so the outcome for a given u is reproducible; atomics only order shots inside a bucket.
constexpr int kLogB = 12;
constexpr long long kB = 1ll << kLogB;
constexpr int kSliceShots = 1024; // shots one CTA of pass 2 serves; a block with more gets extra CTAs from a device-built list

__device__ __forceinline__ double4 ld_nc_256(const double *p) { // 32-byte aligned, read once
    double4 v;
    asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}
// One atomic per warp and distinct key instead of one per lane: a peaked distribution sends most of 2^20 shots to a handful of
// blocks, and a million atomics on one address serialise.  Returns this lane's slot: the old counter value plus its rank among
// the lanes of the warp that hold the same key.
__device__ __forceinline__ int warp_slot(int *counter, int key, bool active) {
    const unsigned live = __ballot_sync(0xffffffffu, active);
    if (!active) return 0;
    const unsigned same = __match_any_sync(live, key);
    const int leader = __ffs(same) - 1, lane = threadIdx.x & 31;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter + key, __popc(same));
    base = __shfl_sync(same, base, leader);
    return base + __popc(same & ((1u << lane) - 1u));
}

// out[b] = sum over i in [b*B, (b+1)*B) of |amp_i|^2 ; one CTA per block, fixed summation order.  Full blocks are read as 256-bit
// vectors, all eight loads of a thread in flight before the first add (the pass is a pure HBM read).
__global__ void __launch_bounds__(256) k_block_prob(const double *__restrict__ re, const double *__restrict__ im,
                                                    long long len, double *__restrict__ out) {
    const long long b = blockIdx.x;
    const long long lo = b * kB, hi = min(lo + kB, len);
    double acc = 0.0;
    if (hi - lo == kB) {
        double4 x[4], y[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            x[i] = ld_nc_256(re + lo + 4 * (i * 256 + threadIdx.x));
            y[i] = ld_nc_256(im + lo + 4 * (i * 256 + threadIdx.x));
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            acc += (x[i].x * x[i].x + y[i].x * y[i].x) + (x[i].y * x[i].y + y[i].y * y[i].y) + (x[i].z * x[i].z + y[i].z * y[i].z) +
                   (x[i].w * x[i].w + y[i].w * y[i].w);
    } else {
        for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) acc += re[i] * re[i] + im[i] * im[i];
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0) out[b] = acc;
}

// Inclusive scan of 4096 values held 16 per thread (thread t owns v[16t .. 16t+16)) by a 256-thread CTA, in place.
// Returns the total in every thread.
template <typename T>
__device__ __forceinline__ T cta_scan16(T (&v)[16]) {
    __shared__ T warp_tot[8];
    __shared__ T cta_tot;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 1; i < 16; ++i) v[i] += v[i - 1];
    T incl = v[15];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    T base = incl - v[15];
    for (int k = 0; k < w; ++k) base += warp_tot[k];
    if (threadIdx.x == 255) cta_tot = base + v[15];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += base;
    __syncthreads();
    const T total = cta_tot;
    __syncthreads(); // the statics are reused by the caller's next scan
    return total;
}

// One CTA per group of 4096 entries: local inclusive prefix (in place allowed) and the group's total.
template <typename T>
__global__ void __launch_bounds__(256) k_scan_groups(const T *__restrict__ in, long long n_in, T *__restrict__ local_incl, T *__restrict__ totals) {
    const long long lo = (long long)blockIdx.x * kB;
    T v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const long long j = lo + threadIdx.x * 16 + i;
        v[i] = j < n_in ? in[j] : (T)0;
    }
    const T total = cta_scan16(v);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const long long j = lo + threadIdx.x * 16 + i;
        if (j < n_in) local_incl[j] = v[i];
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = total;
}
// Inclusive scan of up to 4096 group totals, in place, one CTA.
template <typename T>
__global__ void __launch_bounds__(256) k_scan_single(T *__restrict__ x, int n) {
    T v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int j = threadIdx.x * 16 + i;
        v[i] = j < n ? x[j] : (T)0;
    }
    cta_scan16(v);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int j = threadIdx.x * 16 + i;
        if (j < n) x[j] = v[i];
    }
}

// first j in [0, cnt) with a[j] > x (cnt - 1 if rounding leaves none)
__device__ __forceinline__ long long upper_index(const double *__restrict__ a, long long cnt, double x) {
    long long lo = 0, hi = cnt;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (a[mid] > x) hi = mid; else lo = mid + 1;
    }
    return lo < cnt ? lo : cnt - 1;
}

// One thread per shot: the block that holds it, the residual inside the block, and the histogram of shots per block.
__global__ void __launch_bounds__(256) k_locate(const double *__restrict__ u01, long long shots, const double *__restrict__ pre1, long long n_l1,
                                                const double *__restrict__ pre2, int n_l2, int *__restrict__ shot_blk,
                                                double *__restrict__ shot_res, int *__restrict__ count) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = s < shots;
    int blk = 0;
    if (active) {
        double u = u01[s];
        if (!(u >= 0.0)) u = 0.0;
        if (u >= 1.0) u = 0x1.fffffffffffffp-1;
        const double x = u * pre2[n_l2 - 1];
        const long long g = upper_index(pre2, n_l2, x);
        const double r1 = x - (g ? pre2[g - 1] : 0.0);
        const long long lo = g * kB, cnt = min(kB, n_l1 - lo);
        const long long j = upper_index(pre1 + lo, cnt, r1);
        const double r2 = r1 - (j ? pre1[lo + j - 1] : 0.0);
        blk = (int)(lo + j);
        shot_blk[s] = blk;
        shot_res[s] = r2;
    }
    warp_slot(count, blk, active);
}
// bucket start of block b from the two-level inclusive scan of the histogram
__device__ __forceinline__ int bucket_start(const int *__restrict__ count, const int *__restrict__ cnt_incl, const int *__restrict__ cnt_tot, long long b) {
    const long long g = b >> kLogB;
    return cnt_incl[b] - count[b] + (g ? cnt_tot[g - 1] : 0);
}
__global__ void __launch_bounds__(256) k_scatter(long long shots, const int *__restrict__ shot_blk, const int *__restrict__ count,
                                                 const int *__restrict__ cnt_incl, const int *__restrict__ cnt_tot, int *__restrict__ cursor,
                                                 int *__restrict__ order) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = s < shots;
    const int b = active ? shot_blk[s] : 0;
    const int slot = warp_slot(cursor, b, active);
    if (active) order[bucket_start(count, cnt_incl, cnt_tot, b) + slot] = (int)s;
}
// The extra work of pass 2: block b holds count[b] shots, its own CTA serves the first kSliceShots of them, and every further
// slice of kSliceShots becomes an item (b, slice) of a list -- at most shots / kSliceShots items in all.
__global__ void __launch_bounds__(256) k_hot_list(const int *__restrict__ count, long long n_l1, int *__restrict__ n_items, int2 *__restrict__ items,
                                                  int max_items) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_l1) return;
    const int extra = (count[b] - 1) / kSliceShots; // slices 1 .. extra
    if (extra <= 0) return;
    const int at = atomicAdd(n_items, extra);
    for (int k = 0; k < extra && at + k < max_items; ++k) items[at + k] = make_int2((int)b, k + 1);
}
// One CTA per block with shots: scan the block's probabilities once, answer the shots [slice * kSliceShots, (slice + 1) *
// kSliceShots) of its bucket.  Launched twice: over all blocks (slice 0; a block without shots is not read), and over the
// device-built list of further slices of the blocks that hold more than kSliceShots shots (items != nullptr).
__global__ void __launch_bounds__(256) k_sample_blocks(const double *__restrict__ re, const double *__restrict__ im, long long len,
                                                       const int *__restrict__ count, const int *__restrict__ cnt_incl,
                                                       const int *__restrict__ cnt_tot, const int *__restrict__ order,
                                                       const double *__restrict__ shot_res, long long *__restrict__ out,
                                                       const int *__restrict__ n_items, const int2 *__restrict__ items) {
    long long b = blockIdx.x;
    int slice = 0;
    if (items) {
        if ((int)blockIdx.x >= *n_items) return;
        const int2 it = items[blockIdx.x];
        b = it.x; slice = it.y;
    }
    const int n_here = count[b];
    if (n_here <= slice * kSliceShots) return;
    __shared__ double incl[kB];
    const long long lo = b * kB, cnt = min(kB, len - lo);
    double v[16];
    if (cnt == kB) { // 256-bit reads, parked in shared memory in index order
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int j = 4 * (i * 256 + threadIdx.x);
            const double4 x = ld_nc_256(re + lo + j), y = ld_nc_256(im + lo + j);
            incl[j] = x.x * x.x + y.x * y.x; incl[j + 1] = x.y * x.y + y.y * y.y;
            incl[j + 2] = x.z * x.z + y.z * y.z; incl[j + 3] = x.w * x.w + y.w * y.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const long long j = (long long)i * 256 + threadIdx.x;
            double p = 0.0;
            if (j < cnt) { const double x = re[lo + j], y = im[lo + j]; p = x * x + y * y; }
            incl[j] = p;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = incl[threadIdx.x * 16 + i];
    __syncthreads();
    cta_scan16(v);
#pragma unroll
    for (int i = 0; i < 16; ++i) incl[threadIdx.x * 16 + i] = v[i];
    __syncthreads();
    const int start = bucket_start(count, cnt_incl, cnt_tot, b);
    const int k_end = min(n_here, (slice + 1) * kSliceShots);
    for (int k = slice * kSliceShots + threadIdx.x; k < k_end; k += blockDim.x) {
        const int s = order[start + k];
        out[s] = lo + upper_index(incl, cnt, shot_res[s]);
    }
}

// The sampler's device scratch: one stream-ordered arena kept with the state and grown on demand (cudaMalloc / cudaFree would
// synchronise the whole device, see spz_create).
static int sample_arena(spz_state *st, size_t bytes, char **out) {
    if (st->scratch.samp_bytes < bytes) {
        if (st->scratch.samp) SPZ_CUDA(cudaFreeAsync(st->scratch.samp, st->stream));
        st->scratch.samp = nullptr; st->scratch.samp_bytes = 0;
        const size_t cap = bytes + bytes / 4;
        SPZ_CUDA(cudaMallocAsync(&st->scratch.samp, cap, st->stream));
        st->scratch.samp_bytes = cap;
    }
    *out = static_cast<char *>(st->scratch.samp);
    return SPZ_OK;
}

int launch_sample(spz_state *st, const double *u01, int64_t shots, int64_t *out_index) {
    if (shots <= 0) return SPZ_OK;
    if (shots > 0x7fffffffLL) { set_error("at most 2^31 - 1 shots per call"); return SPZ_ERR_INVALID_ARG; }
    SPZ_TRY(join_pending(st));
    const long long len = st->len;
    const long long n_l1 = (len + kB - 1) / kB;
    const int n_l2 = (int)((n_l1 + kB - 1) / kB);
    if (n_l2 > (int)kB) { set_error("register too large for the two-level sampling CDF"); return SPZ_ERR_UNSUPPORTED; }
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_pre1 = 0, o_pre2 = o_pre1 + up(8 * (size_t)n_l1), o_count = o_pre2 + up(8 * (size_t)n_l2);
    const size_t o_cursor = o_count + up(4 * (size_t)n_l1), o_cincl = o_cursor + up(4 * (size_t)n_l1), o_ctot = o_cincl + up(4 * (size_t)n_l1);
    const size_t o_u = o_ctot + up(4 * (size_t)n_l2), o_res = o_u + up(8 * (size_t)shots), o_blk = o_res + up(8 * (size_t)shots);
    const int max_items = (int)(shots / kSliceShots) + 1;
    const size_t o_order = o_blk + up(4 * (size_t)shots), o_out = o_order + up(4 * (size_t)shots), o_items = o_out + up(8 * (size_t)shots);
    const size_t o_nitems = o_items + up(8 * (size_t)max_items), total = o_nitems + 256;
    char *base = nullptr;
    SPZ_TRY(sample_arena(st, total, &base));
    double *pre1 = reinterpret_cast<double *>(base + o_pre1), *pre2 = reinterpret_cast<double *>(base + o_pre2);
    int *count = reinterpret_cast<int *>(base + o_count), *cursor = reinterpret_cast<int *>(base + o_cursor);
    int *cincl = reinterpret_cast<int *>(base + o_cincl), *ctot = reinterpret_cast<int *>(base + o_ctot);
    double *d_u = reinterpret_cast<double *>(base + o_u), *d_res = reinterpret_cast<double *>(base + o_res);
    int *d_blk = reinterpret_cast<int *>(base + o_blk), *d_order = reinterpret_cast<int *>(base + o_order);
    long long *d_out = reinterpret_cast<long long *>(base + o_out);
    int2 *d_items = reinterpret_cast<int2 *>(base + o_items);
    int *d_nitems = reinterpret_cast<int *>(base + o_nitems);
    cudaStream_t q = st->stream;
    SPZ_CUDA(cudaMemsetAsync(d_nitems, 0, sizeof(int), q));
    SPZ_CUDA(cudaMemcpyAsync(d_u, u01, sizeof(double) * (size_t)shots, cudaMemcpyHostToDevice, q));
    SPZ_CUDA(cudaMemsetAsync(count, 0, o_cincl - o_count, q)); // histogram and cursors
    k_block_prob<<<(unsigned)n_l1, 256, 0, q>>>(st->re, st->im, len, pre1);
    k_scan_groups<double><<<(unsigned)n_l2, 256, 0, q>>>(pre1, n_l1, pre1, pre2);
    k_scan_single<double><<<1, 256, 0, q>>>(pre2, n_l2);
    const unsigned shot_grid = (unsigned)((shots + 255) / 256);
    k_locate<<<shot_grid, 256, 0, q>>>(d_u, shots, pre1, n_l1, pre2, n_l2, d_blk, d_res, count);
    k_scan_groups<int><<<(unsigned)n_l2, 256, 0, q>>>(count, n_l1, cincl, ctot);
    k_scan_single<int><<<1, 256, 0, q>>>(ctot, n_l2);
    k_scatter<<<shot_grid, 256, 0, q>>>(shots, d_blk, count, cincl, ctot, cursor, d_order);
    k_hot_list<<<(unsigned)((n_l1 + 255) / 256), 256, 0, q>>>(count, n_l1, d_nitems, d_items, max_items);
    k_sample_blocks<<<(unsigned)n_l1, 256, 0, q>>>(st->re, st->im, len, count, cincl, ctot, d_order, d_res, d_out, nullptr, nullptr);
    k_sample_blocks<<<(unsigned)max_items, 256, 0, q>>>(st->re, st->im, len, count, cincl, ctot, d_order, d_res, d_out, d_nitems, d_items);
    count_launch(10);
    SPZ_CUDA(cudaGetLastError());
    SPZ_CUDA(cudaMemcpyAsync(out_index, d_out, sizeof(long long) * (size_t)shots, cudaMemcpyDeviceToHost, q));
    SPZ_CUDA(cudaStreamSynchronize(q));
    return SPZ_OK;
}

} // namespace spz
