// dist_plan.h -- host-side lowering of logical gates onto a sharded register (no CUDA in this file).
//
// The 2^n amplitudes are sharded over P = 2^g ranks by the top g PHYSICAL index bits: rank r owns
// physical indices [r * 2^(n-g), (r+1) * 2^(n-g)).  A permutation perm[logical qubit] = physical bit is kept
// instead of moving qubits back: physical bits 0..n-g-1 are local, n-g..n-1 are the bits of the rank.
// The reference has no distributed layer at all; this is new design.
//
//   local target                        -> the single-GPU kernels on the shard, no communication
//   diagonal gate (Z/P/RZ), any qubits  -> never communicates: a global control/target bit is a per-rank
//                                          constant (skip / constant factor on the shard)
//   non-diagonal gate, global target    -> EXCHANGE: swap the global physical bit with a local one (pairwise
//                                          half-shard exchange with rank ^ (1 << k), dist.cu), update perm, then
//                                          the gate is local
//   SWAP(a, b)                          -> relabel perm[a] <-> perm[b]; no data moves
//
// The plan is identical on every rank (SPMD): `rank` only decides skip / which constant.  The victim local
// qubit of an EXCHANGE is the resident qubit whose next non-diagonal use is farthest away when execute()
// provides the op list (Belady).  Gate-by-gate calls have no look-ahead; there the victim is the MOST recently
// used resident qubit: circuits sweep over qubits layer by layer (and QFT never touches a qubit non-diagonally
// again after its H), so the qubit just used is the one needed latest -- LRU would evict exactly the qubit the
// sweep needs next and thrash (measured: 254 exchanges instead of 12 in the 2-GPU bandwidth sweep).
// Candidates are physical bits >= kMinExchangeBit so that exchanged runs are long contiguous segments.
#pragma once

#include <stdint.h>

#include <algorithm>
#include <vector>

#include "../../include/spinoza_b200.h"

struct spz_state;

#define SPZ_TRY_PLAN(x) do { int rc__ = (x); if (rc__ != SPZ_OK) return rc__; } while (0)

namespace spz {

enum { ACT_SKIP = 0, ACT_LOCAL_GATE = 1, ACT_DIAG_CONST = 2, ACT_EXCHANGE = 3, ACT_LOCAL_SWAP = 4 };

struct DistPlan {
    int n = 0, g = 0, n_local = 0, world = 1;
    int perm[64];          // logical -> physical
    int inv[64];           // physical -> logical
    uint64_t last_use[64]; // LRU clock per logical qubit
    uint64_t clock = 1;
    static constexpr int kMinExchangeBit = 8;

    void init(int n_total, int world_size) {
        n = n_total; world = world_size; g = 0;
        while ((1 << g) < world_size) ++g;
        n_local = n - g;
        for (int q = 0; q < 64; ++q) { perm[q] = inv[q] = q; last_use[q] = 0; }
    }
    bool is_global_phys(int p) const { return p >= n_local; }
    void touch(int logical) { last_use[logical] = clock++; }

    // next_use[logical] = distance to the next non-diagonal use (UINT64_MAX = never), or nullptr -> LRU
    int choose_victim(const uint64_t *next_use, uint64_t protect_mask) const {
        int best = -1;
        const int lo = std::min(kMinExchangeBit, std::max(0, n_local - 1));
        for (int p = n_local - 1; p >= lo; --p) {
            const int l = inv[p];
            if ((protect_mask >> l) & 1ull) continue;
            if (best < 0) { best = p; continue; }
            const int bl = inv[best];
            if (next_use) {
                if (next_use[l] > next_use[bl]) best = p; // Belady: farthest next non-diagonal use
            } else if (last_use[l] > last_use[bl]) best = p; // no look-ahead: MOST recently used (see header)
        }
        if (best < 0) // every candidate is protected: fall back to any unprotected local bit
            for (int p = n_local - 1; p >= 0; --p)
                if (!((protect_mask >> inv[p]) & 1ull)) { best = p; break; }
        return best;
    }

    // Make logical qubit q resident (physical bit < n_local).  Appends at most one EXCHANGE action.
    int ensure_local(int q, const uint64_t *next_use, uint64_t protect_mask, std::vector<spz_dist_action> &out, int rank) {
        const int p = perm[q];
        if (!is_global_phys(p)) return SPZ_OK;
        // prefer not to evict a control of this gate, but any resident qubit will do
        int victim_p = choose_victim(next_use, protect_mask | (1ull << q));
        if (victim_p < 0) victim_p = choose_victim(next_use, 1ull << q);
        if (victim_p < 0) return SPZ_ERR_INVALID_ARG;
        const int victim_l = inv[victim_p];
        spz_dist_action a{};
        a.type = ACT_EXCHANGE;
        a.gbit = p - n_local;          // which bit of the rank
        a.lq = victim_p;               // local physical bit it trades places with
        a.partner = rank ^ (1 << a.gbit);
        out.push_back(a);
        perm[q] = victim_p; inv[victim_p] = q;
        perm[victim_l] = p; inv[p] = victim_l;
        return SPZ_OK;
    }

    // Lower one logical op.  kind: spz_gate_kind (1-qubit kinds or SWAP); cmask: logical all-ones controls.
    int lower(int rank, int kind, const double *params, int t0, int t1, uint64_t cmask, int target,
              const uint64_t *next_use, std::vector<spz_dist_action> &out) {
        if (kind == SPZ_GATE_SWAP) {
            if (t0 < 0 || t1 < 0 || t0 >= n || t1 >= n) return SPZ_ERR_INVALID_ARG;
            if (t0 == t1) return SPZ_OK;
            const int p0 = perm[t0], p1 = perm[t1];
            perm[t0] = p1; perm[t1] = p0; inv[p1] = t0; inv[p0] = t1; // relabel: no data moves
            touch(t0); touch(t1);
            return SPZ_OK;
        }
        if (target < 0 || target >= n || ((cmask >> target) & 1ull) || (n < 64 && (cmask >> n))) return SPZ_ERR_INVALID_ARG;
        const bool diag = kind == SPZ_GATE_Z || kind == SPZ_GATE_P || kind == SPZ_GATE_RZ;
        if (!diag) SPZ_TRY_PLAN(ensure_local(target, next_use, cmask, out, rank));
        touch(target);
        // resolve controls: local -> mask bit, global -> per-rank constant
        uint64_t local_cmask = 0;
        bool skip = false;
        int grefs = 0; // rank bits this lowering depends on (the same on every rank, whatever their values)
        for (int c = 0; c < n; ++c) {
            if (!((cmask >> c) & 1ull)) continue;
            const int pc = perm[c];
            if (is_global_phys(pc)) { grefs |= 1 << (pc - n_local); if (!((rank >> (pc - n_local)) & 1)) skip = true; }
            else local_cmask |= 1ull << pc;
        }
        spz_dist_action a{};
        a.kind = kind;
        if (params) { a.p[0] = params[0]; a.p[1] = params[1]; a.p[2] = params[2]; }
        a.cmask = local_cmask;
        const int pt = perm[target];
        // A skipped op is still described completely (target, controls, grefs): a scheduler that wants the same pass
        // structure on every rank treats it as a placeholder with the dependencies of the real op.
        if (!is_global_phys(pt)) {
            a.type = ACT_LOCAL_GATE; a.target = pt;
        } else { // diagonal gate on a global target: constant per rank
            grefs |= 1 << (pt - n_local);
            a.type = ACT_DIAG_CONST; a.target = -1; a.hi = (rank >> (pt - n_local)) & 1;
            if (!a.hi && kind != SPZ_GATE_RZ) skip = true; // Z / P leave target-bit-0 amplitudes alone
        }
        a.grefs = grefs;
        if (skip) a.type = ACT_SKIP;
        out.push_back(a);
        return SPZ_OK;
    }
};

// implemented in dist.cu; declared here because it needs spz_dist_action
int dist_lower(struct ::spz_state *st, int kind, const double *p, int t0, int t1, uint64_t cmask, int target,
               const uint64_t *next_use, std::vector<spz_dist_action> &acts);

} // namespace spz
