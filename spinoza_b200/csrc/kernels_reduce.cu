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

namespace spz {

constexpr int kRedThreads = 256;
constexpr int kRedBlocks = 148 * 8; // grid sized in multiples of the SM count

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
        SPZ_CUDA(cudaMalloc(&st->scratch.partials, sizeof(double) * (kRedBlocks + 8)));
        st->scratch.n_partials = kRedBlocks + 8;
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
// Exact inverse-CDF sampling with a three-level CDF (replaces the O(num_tests * k) reservoir loop of
// core.rs:81-112).  Level 0: |amp|^2.  Level 1: sums of blocks of B amplitudes.  Level 2: sums of groups of
// B level-1 entries.  A shot x = u * total is located by a host search over level 2 (<= 2^(n-2*LOGB)
// entries), then one warp per shot scans its level-1 group and its amplitude block.
constexpr int kLogB = 12;
constexpr long long kB = 1ll << kLogB;

// out[b] = sum over i in [b*B, (b+1)*B) of |amp_i|^2 ; one CTA per block, fixed summation order
__global__ void __launch_bounds__(256) k_block_prob(const double *__restrict__ re, const double *__restrict__ im,
                                                    long long len, double *__restrict__ out) {
    const long long b = blockIdx.x;
    const long long lo = b * kB, hi = min(lo + kB, len);
    double acc = 0.0;
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) acc += re[i] * re[i] + im[i] * im[i];
    acc = block_sum(acc);
    if (threadIdx.x == 0) out[b] = acc;
}
__global__ void __launch_bounds__(256) k_group_sum(const double *__restrict__ in, long long n_in, double *__restrict__ out) {
    const long long g = blockIdx.x;
    const long long lo = g * kB, hi = min(lo + kB, n_in);
    double acc = 0.0;
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) acc += in[i];
    acc = block_sum(acc);
    if (threadIdx.x == 0) out[g] = acc;
}

// Warp-cooperative search: smallest j in [0, cnt) with prefix(j) > x, where prefix accumulates w(lo + j) in a
// fixed order (lane-contiguous chunks, then lanes in order).  Returns j (clamped to cnt-1) and leaves the
// residual x - prefix(j-1) in *resid.
template <typename F>
__device__ __forceinline__ long long warp_search(F w, long long cnt, double x, double *resid) {
    const int lane = threadIdx.x & 31;
    const long long per = (cnt + 31) / 32;
    const long long b = min((long long)lane * per, cnt), e = min(b + per, cnt);
    double mine = 0.0;
    for (long long j = b; j < e; ++j) mine += w(j);
    // exclusive prefix over lanes
    double incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const double excl = incl - mine;
    // owning lane: first lane with incl > x (last non-empty lane if rounding leaves none)
    const unsigned ballot = __ballot_sync(0xffffffffu, incl > x && e > b);
    int owner;
    if (ballot) owner = __ffs(ballot) - 1;
    else {
        const unsigned nonempty = __ballot_sync(0xffffffffu, e > b);
        owner = 31 - __clz(nonempty);
    }
    long long found = 0;
    double r = 0.0;
    if (lane == owner) {
        double acc = excl;
        long long j = b;
        for (; j < e; ++j) {
            const double wj = w(j);
            if (acc + wj > x) break;
            acc += wj;
        }
        if (j >= e) j = e - 1; // rounding guard
        found = j;
        // residual relative to the start of element j
        double acc2 = excl;
        for (long long q = b; q < j; ++q) acc2 += w(q);
        r = x - acc2;
    }
    found = __shfl_sync(0xffffffffu, found, owner);
    r = __shfl_sync(0xffffffffu, r, owner);
    *resid = r;
    return found;
}

__global__ void __launch_bounds__(256) k_sample(const double *__restrict__ re, const double *__restrict__ im, long long len,
                                                const double *__restrict__ l1, long long n_l1,
                                                const long long *__restrict__ shot_group, const double *__restrict__ shot_resid,
                                                long long shots, long long *__restrict__ out) {
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long s = warp; s < shots; s += nwarps) {
        const long long g = shot_group[s];
        double x = shot_resid[s];
        const long long l1_lo = g * kB, l1_cnt = min(kB, n_l1 - l1_lo);
        double r1;
        const long long jb = warp_search([&](long long j) { return l1[l1_lo + j]; }, l1_cnt, x, &r1);
        const long long blk = l1_lo + jb;
        const long long a_lo = blk * kB, a_cnt = min(kB, len - a_lo);
        double r2;
        const long long ja = warp_search(
            [&](long long j) { const double a = re[a_lo + j], b = im[a_lo + j]; return a * a + b * b; }, a_cnt, r1, &r2);
        if ((threadIdx.x & 31) == 0) out[s] = a_lo + ja;
    }
}

int launch_sample(spz_state *st, const double *u01, int64_t shots, int64_t *out_index) {
    if (shots <= 0) return SPZ_OK;
    SPZ_TRY(join_pending(st));
    const long long len = st->len;
    const long long n_l1 = (len + kB - 1) / kB;
    const long long n_l2 = (n_l1 + kB - 1) / kB;
    double *d_l1 = nullptr, *d_l2 = nullptr, *d_resid = nullptr;
    long long *d_group = nullptr, *d_out = nullptr;
    int rc = SPZ_OK;
    std::vector<double> l2((size_t)n_l2), resid((size_t)shots);
    std::vector<long long> group((size_t)shots);
    // stream-ordered allocations: cudaMalloc/cudaFree would synchronise the whole device (see spz_create)
    auto cleanup = [&]() {
        cudaFreeAsync(d_l1, st->stream); cudaFreeAsync(d_l2, st->stream); cudaFreeAsync(d_resid, st->stream);
        cudaFreeAsync(d_group, st->stream); cudaFreeAsync(d_out, st->stream);
        cudaStreamSynchronize(st->stream);
    };
#define SPZ_S(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { rc = cuda_fail(e__, #call, __FILE__, __LINE__); cleanup(); return rc; } } while (0)
    SPZ_S(cudaMallocAsync(&d_l1, sizeof(double) * (size_t)n_l1, st->stream));
    SPZ_S(cudaMallocAsync(&d_l2, sizeof(double) * (size_t)n_l2, st->stream));
    SPZ_S(cudaMallocAsync(&d_resid, sizeof(double) * (size_t)shots, st->stream));
    SPZ_S(cudaMallocAsync(&d_group, sizeof(long long) * (size_t)shots, st->stream));
    SPZ_S(cudaMallocAsync(&d_out, sizeof(long long) * (size_t)shots, st->stream));
    k_block_prob<<<(unsigned)n_l1, 256, 0, st->stream>>>(st->re, st->im, len, d_l1);
    k_group_sum<<<(unsigned)n_l2, 256, 0, st->stream>>>(d_l1, n_l1, d_l2);
    count_launch(2);
    SPZ_S(cudaGetLastError());
    SPZ_S(cudaMemcpyAsync(l2.data(), d_l2, sizeof(double) * (size_t)n_l2, cudaMemcpyDeviceToHost, st->stream));
    SPZ_S(cudaStreamSynchronize(st->stream));
    // host: level-2 CDF and per-shot group + residual
    std::vector<double> cdf2((size_t)n_l2);
    double total = 0.0;
    for (long long g = 0; g < n_l2; ++g) { total += l2[(size_t)g]; cdf2[(size_t)g] = total; }
    for (int64_t s = 0; s < shots; ++s) {
        double u = u01[s];
        if (!(u >= 0.0)) u = 0.0;
        if (u >= 1.0) u = 0x1.fffffffffffffp-1;
        const double x = u * total;
        long long g = std::upper_bound(cdf2.begin(), cdf2.end(), x) - cdf2.begin();
        if (g >= n_l2) g = n_l2 - 1;
        group[(size_t)s] = g;
        resid[(size_t)s] = x - (g ? cdf2[(size_t)g - 1] : 0.0);
    }
    SPZ_S(cudaMemcpyAsync(d_group, group.data(), sizeof(long long) * (size_t)shots, cudaMemcpyHostToDevice, st->stream));
    SPZ_S(cudaMemcpyAsync(d_resid, resid.data(), sizeof(double) * (size_t)shots, cudaMemcpyHostToDevice, st->stream));
    const long long warps_needed = shots;
    const int grid = (int)std::max<long long>(1, std::min<long long>((warps_needed * 32 + 255) / 256, 148 * 16));
    k_sample<<<grid, 256, 0, st->stream>>>(st->re, st->im, len, d_l1, n_l1, d_group, d_resid, shots, d_out);
    count_launch();
    SPZ_S(cudaGetLastError());
    SPZ_S(cudaMemcpyAsync(out_index, d_out, sizeof(long long) * (size_t)shots, cudaMemcpyDeviceToHost, st->stream));
    SPZ_S(cudaStreamSynchronize(st->stream));
#undef SPZ_S
    cleanup();
    return SPZ_OK;
}

} // namespace spz
