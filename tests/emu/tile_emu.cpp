// tile_emu.cpp -- runs the bodies of the fused tile kernels k_tile (spinoza_b200/csrc/kernels_tile.cu) and k_tile3
// (kernels_tile3.cu: the TMA kernel; its box copies become synchronous swizzled copies here) on the CPU, block by block.
//
// Test infrastructure only: built by tests/test_tile_cpu_emulation.py with
//   g++ -O1 -std=c++17 -ffp-contract=off -shared -fPIC -pthread -I/usr/local/cuda/include -include tests/emu/cuda_cpu_shim.h
// The kernel sources are #included unchanged; see cuda_cpu_shim.h for what the emulation does and does not model.
#define SPZ_CPU_EMULATION 1
#include "cuda_cpu_shim.h"

#include <cstring>
#include <thread>
#include <vector>

#include "../../spinoza_b200/csrc/kernels_tile.cu"
#include "../../spinoza_b200/csrc/kernels_tile3.cu"

namespace spz_emu {
unsigned char *dyn_smem = nullptr;
CtaBarrier default_cta;
} // namespace spz_emu

namespace {

// One OS thread per CUDA thread; the blocks of the grid run one after another.
template <class Kernel>
void run_grid(unsigned n_blocks, unsigned n_threads, size_t smem_bytes, Kernel kernel) {
    std::vector<unsigned char> window(smem_bytes + 64);
    spz_emu::dyn_smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(window.data()) + 63) & ~(uintptr_t)63);
    spz_emu::default_cta.init(n_threads);
    std::vector<std::thread> pool;
    pool.reserve(n_threads);
    for (unsigned t = 0; t < n_threads; ++t) {
        pool.emplace_back([&, t]() {
            threadIdx.x = t; threadIdx.y = 0; threadIdx.z = 0;
            blockDim.x = n_threads; blockDim.y = 1; blockDim.z = 1;
            gridDim.x = n_blocks; gridDim.y = 1; gridDim.z = 1;
            for (unsigned b = 0; b < n_blocks; ++b) {
                blockIdx.x = b; blockIdx.y = 0; blockIdx.z = 0;
                kernel();
                spz_emu::barrier(); // the next block reuses the shared-memory statics
            }
        });
    }
    for (auto &th : pool) th.join();
    spz_emu::default_cta.destroy();
    spz_emu::dyn_smem = nullptr;
}

struct Program {
    spz::TilePlan plan{};
    std::vector<spz::TileInstr> prog;
    std::vector<spz::TileGroup> groups;
    std::vector<spz::TileTerm> terms;
    int ni = 0, ng = 0, nt = 0;
};

// blob: the serialisation produced by spz_debug_compile_pass (layout documented in csrc/abi.cu)
int parse(const void *blob, long long blob_bytes, Program &P) {
    const char *p = static_cast<const char *>(blob);
    int32_t hdr[16];
    if (blob_bytes < (long long)sizeof hdr) return -1;
    std::memcpy(hdr, p, sizeof hdr);
    if (hdr[0] != 0 || hdr[15] != (int32_t)sizeof(spz::TileInstr)) return -2;
    P.plan.tile_bits = hdr[1]; P.plan.low_bits = hdr[2]; P.plan.n_high = hdr[3];
    for (int k = 0; k < 8; ++k) P.plan.high[k] = hdr[4 + k];
    P.ni = hdr[12]; P.ng = hdr[13]; P.nt = hdr[14];
    P.prog.resize(P.ni);
    P.groups.resize(P.ng > 0 ? P.ng : 1);
    P.terms.resize(P.nt > 0 ? P.nt : 1);
    p += sizeof hdr;
    std::memcpy(P.prog.data(), p, sizeof(spz::TileInstr) * P.ni); p += sizeof(spz::TileInstr) * P.ni;
    if (P.ng) std::memcpy(P.groups.data(), p, sizeof(spz::TileGroup) * P.ng);
    p += sizeof(spz::TileGroup) * P.ng;
    if (P.nt) std::memcpy(P.terms.data(), p, sizeof(spz::TileTerm) * P.nt);
    return 0;
}

} // namespace

// k_tile3, through the host lowering the launcher uses (tile3_lower / tile3_pack).  info[0..3] <- {some butterfly has an
// in-tile control, instructions, per-tile groups, shared-memory bytes}; returns 1 when the launcher would fall back to k_tile.
extern "C" int emu_tile3_run(int n_qubits, double *re, double *im, const void *blob, long long blob_bytes, int *info) {
    Program P;
    if (int rc = parse(blob, blob_bytes, P)) return rc;
    if (!spz::tile3_shape_ok(n_qubits, P.plan)) return 1;
    spz::Lowered3 lw;
    if (!spz::tile3_lower(P.plan, P.prog.data(), P.ni, P.groups.data(), P.ng, P.terms.data(), P.nt, lw)) return 1;
    spz::Tile3Args a{};
    std::vector<unsigned char> packed;
    spz::tile3_pack(lw, P.plan, packed, a);
    const size_t smem = spz::tile3_smem_bytes(a);
    if (smem > spz::kSmemBudget3) return 1;
    a.re = re; a.im = im;
    a.blob = packed.data();
    info[0] = lw.ctrl ? 1 : 0; info[1] = a.n_ins; info[2] = a.n_groups; info[3] = (int)smem;
    const unsigned n_tiles = (unsigned)(((uint64_t)1 << n_qubits) >> P.plan.tile_bits);
    a.tile_first = 0; a.tile_end = n_tiles;
    // the kernel is persistent: fewer CTAs than tiles, so that every CTA walks several (the launcher uses two per SM)
    run_grid(n_tiles > 3 ? 3 : n_tiles, spz::kThreads3, smem, [&]() { spz::k_tile3(a); });
    return 0;
}

// Shared-memory bytes k_tile3 needs for this pass (0: the lowering refuses it), without running anything.  info as above.
extern "C" long long emu_tile3_smem(int n_qubits, const void *blob, long long blob_bytes, int *info) {
    Program P;
    if (parse(blob, blob_bytes, P)) return -1;
    if (!spz::tile3_shape_ok(n_qubits, P.plan)) return 0;
    spz::Lowered3 lw;
    if (!spz::tile3_lower(P.plan, P.prog.data(), P.ni, P.groups.data(), P.ng, P.terms.data(), P.nt, lw)) return 0;
    spz::Tile3Args a{};
    std::vector<unsigned char> packed;
    spz::tile3_pack(lw, P.plan, packed, a);
    info[0] = lw.ctrl ? 1 : 0; info[1] = a.n_ins; info[2] = a.n_groups; info[3] = (int)packed.size();
    return (long long)spz::tile3_smem_bytes(a);
}

// The lowered k_tile3 program of one pass, for tools/dump_tile3.py: copies up to max_ins instructions (sizeof(Ins3) = 32 bytes each), returns their
// number (-1: parse error, 0: the lowering refuses the pass).  info <- {in-tile control, instructions, groups, terms}.
extern "C" int emu_tile3_dump(int n_qubits, const void *blob, long long blob_bytes, void *out_ins, int max_ins, int *info) {
    Program P;
    if (parse(blob, blob_bytes, P)) return -1;
    if (!spz::tile3_shape_ok(n_qubits, P.plan)) return 0;
    spz::Lowered3 lw;
    if (!spz::tile3_lower(P.plan, P.prog.data(), P.ni, P.groups.data(), P.ng, P.terms.data(), P.nt, lw)) return 0;
    const int n = (int)lw.ins.size() < max_ins ? (int)lw.ins.size() : max_ins;
    std::memcpy(out_ins, lw.ins.data(), (size_t)n * sizeof(spz::Ins3));
    info[0] = lw.ctrl ? 1 : 0; info[1] = (int)lw.ins.size(); info[2] = (int)lw.groups.size(); info[3] = (int)lw.terms.size();
    return n;
}

// k_tile, with the arguments launch_tile_program (kernels_tile.cu) builds.  prog_in_smem: 0 decodes from "global" memory.
extern "C" int emu_tile1_run(int n_qubits, double *re, double *im, const void *blob, long long blob_bytes, int exact, int prog_in_smem) {
    Program P;
    if (int rc = parse(blob, blob_bytes, P)) return rc;
    const spz::TilePlan &plan = P.plan;
    if (plan.tile_bits > spz::kMaxTileBits || plan.tile_bits < spz::kRegBits || plan.n_high > spz::kMaxHigh || plan.low_bits < 1 ||
        plan.tile_bits != plan.low_bits + plan.n_high || plan.tile_bits > n_qubits)
        return -3;
    spz::TileArgs a{};
    a.re = re; a.im = im;
    a.prog = P.prog.data(); a.groups = P.groups.data(); a.terms = P.terms.data();
    a.n_groups = P.ng; a.n_instr = P.ni;
    a.T = plan.tile_bits; a.L = plan.low_bits; a.n_high = plan.n_high;
    for (int k = 0; k < plan.n_high; ++k) a.high[k] = plan.high[k];
    size_t smem = sizeof(double) * 2u * ((size_t)1 << plan.tile_bits) + (size_t)P.ng * (sizeof(double2) + sizeof(unsigned));
    smem = (smem + 15) & ~(size_t)15;
    a.prog_off = (unsigned)smem;
    a.prog_in_smem = (prog_in_smem && P.ni <= spz::kMaxSmemInstr) ? 1 : 0;
    if (a.prog_in_smem) smem += sizeof(spz::TileInstr) * (size_t)P.ni;
    a.tile_offset = 0;
    const unsigned n_blocks = (unsigned)(((uint64_t)1 << n_qubits) >> plan.tile_bits);
    const unsigned threads = 1u << (plan.tile_bits - spz::kRegBits);
    if (exact) run_grid(n_blocks, threads, smem, [&]() { spz::k_tile<true>(a); });
    else run_grid(n_blocks, threads, smem, [&]() { spz::k_tile<false>(a); });
    return 0;
}

#ifdef SPZ_EMU_MAIN
// Stand-alone driver (used for the ThreadSanitizer run: a sanitised shared object cannot be loaded into CPython):
//   tile_emu <kernel 1|3> <n_qubits> <exact 0|1> <state.bin: re[2^n] then im[2^n], f64> <blob.bin> <option>
// option: k_tile = prog_in_smem 0|1 (ignored by k_tile3).  The state file is rewritten.
#include <cstdio>
#include <cstdlib>
int main(int argc, char **argv) {
    if (argc != 7) return 64;
    const int kernel = std::atoi(argv[1]), n = std::atoi(argv[2]), exact = std::atoi(argv[3]), option = std::atoi(argv[6]);
    const size_t len = (size_t)1 << n;
    std::vector<double> st(2 * len);
    FILE *f = std::fopen(argv[4], "rb");
    if (!f || std::fread(st.data(), sizeof(double), 2 * len, f) != 2 * len) return 65;
    std::fclose(f);
    f = std::fopen(argv[5], "rb");
    if (!f) return 65;
    std::vector<char> blob;
    char buf[1 << 16];
    size_t got;
    while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) blob.insert(blob.end(), buf, buf + got);
    std::fclose(f);
    int info[4] = {0, 0, 0, 0};
    const int rc = kernel == 3 ? emu_tile3_run(n, st.data(), st.data() + len, blob.data(), (long long)blob.size(), info)
                               : emu_tile1_run(n, st.data(), st.data() + len, blob.data(), (long long)blob.size(), exact, option);
    if (rc != 0) return 70 + rc;
    f = std::fopen(argv[4], "wb");
    if (!f || std::fwrite(st.data(), sizeof(double), 2 * len, f) != 2 * len) return 65;
    std::fclose(f);
    std::printf("kernel=%d ctrl=%d instructions=%d groups=%d\n", kernel, info[0], info[1], info[2]);
    return 0;
}
#endif
