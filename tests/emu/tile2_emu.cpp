// tile2_emu.cpp -- runs the body of k_tile2 (spinoza_b200/csrc/kernels_tile2.cu) on the CPU, block by block.
//
// Test infrastructure only: built by tests/test_tile2_cpu_emulation.py with
//   g++ -O1 -std=c++17 -ffp-contract=off -shared -fPIC -pthread -I/usr/local/cuda/include -include tests/emu/cuda_cpu_shim.h
// The kernel source is #included unchanged; see cuda_cpu_shim.h for what the emulation does and does not model.
#define SPZ_CPU_EMULATION 1
#include "cuda_cpu_shim.h"

#include <cstring>
#include <thread>
#include <vector>

#include "../../spinoza_b200/csrc/kernels_tile2.cu"

namespace spz_emu {
unsigned char *dyn_smem = nullptr;
unsigned block_threads = 0;
#ifdef SPZ_EMU_TSAN
std::atomic<unsigned> bar_count{0}, bar_gen{0};
char bar_tags[4096];
#else
pthread_barrier_t block_barrier;
#endif
} // namespace spz_emu

namespace {

template <bool EXACT, bool CTRL>
void run_blocks(const spz::Tile2Args &a, unsigned n_blocks) {
    std::vector<std::thread> pool;
    pool.reserve(spz::kThreads2);
    for (int t = 0; t < spz::kThreads2; ++t) {
        pool.emplace_back([&, t]() {
            threadIdx.x = (unsigned)t; threadIdx.y = 0; threadIdx.z = 0;
            for (unsigned b = 0; b < n_blocks; ++b) {
                blockIdx.x = b; blockIdx.y = 0; blockIdx.z = 0;
                spz::k_tile2<EXACT, CTRL>(a);
                spz_emu::barrier(); // the next block reuses the shared-memory statics
            }
        });
    }
    for (auto &th : pool) th.join();
}

} // namespace

// blob: the serialisation produced by spz_debug_compile_pass (layout documented in csrc/abi.cu).
// info[0..3] <- {ctrl instantiation, first_direct, last_direct, eligible}
extern "C" int emu_tile2_run(int n_qubits, double *re, double *im, const void *blob, long long blob_bytes, int exact, int direct_level,
                             int *info) {
    const char *p = static_cast<const char *>(blob);
    int32_t hdr[16];
    if (blob_bytes < (long long)sizeof hdr) return -1;
    std::memcpy(hdr, p, sizeof hdr);
    if (hdr[0] != 0 || hdr[15] != (int32_t)sizeof(spz::TileInstr)) return -2;
    spz::TilePlan plan{};
    plan.tile_bits = hdr[1]; plan.low_bits = hdr[2]; plan.n_high = hdr[3];
    for (int k = 0; k < 8; ++k) plan.high[k] = hdr[4 + k];
    const int ni = hdr[12], ng = hdr[13], nt = hdr[14];
    std::vector<spz::TileInstr> prog(ni);
    std::vector<spz::TileGroup> groups(ng > 0 ? ng : 1);
    std::vector<spz::TileTerm> terms(nt > 0 ? nt : 1);
    p += sizeof hdr;
    std::memcpy(prog.data(), p, sizeof(spz::TileInstr) * ni); p += sizeof(spz::TileInstr) * ni;
    if (ng) std::memcpy(groups.data(), p, sizeof(spz::TileGroup) * ng);
    p += sizeof(spz::TileGroup) * ng;
    if (nt) std::memcpy(terms.data(), p, sizeof(spz::TileTerm) * nt);
    info[3] = spz::tile2_shape_ok(n_qubits, plan, prog.data(), ni, ng) ? 1 : 0;
    if (!info[3]) return 1; // the launcher would fall back to k_tile
    size_t smem = 0;
    bool ctrl = false;
    const spz::Tile2Args a = spz::tile2_make_args(re, im, plan, prog.data(), ni, prog.data(), groups.data(), ng, terms.data(), 0u, direct_level, &smem, &ctrl);
    info[0] = ctrl ? 1 : 0; info[1] = a.first_direct; info[2] = a.last_direct;
    std::vector<unsigned char> window(smem + 64);
    spz_emu::dyn_smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(window.data()) + 63) & ~(uintptr_t)63);
    spz_emu::block_threads = spz::kThreads2;
#ifndef SPZ_EMU_TSAN
    pthread_barrier_init(&spz_emu::block_barrier, nullptr, spz::kThreads2);
#endif
    const unsigned n_blocks = (unsigned)(((uint64_t)1 << n_qubits) >> plan.tile_bits);
    if (exact) run_blocks<true, true>(a, n_blocks);
    else if (ctrl) run_blocks<false, true>(a, n_blocks);
    else run_blocks<false, false>(a, n_blocks);
#ifndef SPZ_EMU_TSAN
    pthread_barrier_destroy(&spz_emu::block_barrier);
#endif
    spz_emu::dyn_smem = nullptr;
    return 0;
}

#ifdef SPZ_EMU_MAIN
// Stand-alone driver (used for the ThreadSanitizer run: a sanitised shared object cannot be loaded into CPython):
//   tile2_emu <n_qubits> <exact 0|1> <state.bin: re[2^n] then im[2^n], f64> <blob.bin> <direct level 0..3>
// the state file is rewritten.
#include <cstdio>
#include <cstdlib>
int main(int argc, char **argv) {
    if (argc != 6) return 64;
    const int n = std::atoi(argv[1]), exact = std::atoi(argv[2]);
    const size_t len = (size_t)1 << n;
    std::vector<double> st(2 * len);
    FILE *f = std::fopen(argv[3], "rb");
    if (!f || std::fread(st.data(), sizeof(double), 2 * len, f) != 2 * len) return 65;
    std::fclose(f);
    f = std::fopen(argv[4], "rb");
    if (!f) return 65;
    std::vector<char> blob;
    char buf[1 << 16];
    size_t got;
    while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) blob.insert(blob.end(), buf, buf + got);
    std::fclose(f);
    int info[4] = {0, 0, 0, 0};
    const int rc = emu_tile2_run(n, st.data(), st.data() + len, blob.data(), (long long)blob.size(), exact, std::atoi(argv[5]), info);
    if (rc != 0) return 70 + rc;
    f = std::fopen(argv[3], "wb");
    if (!f || std::fwrite(st.data(), sizeof(double), 2 * len, f) != 2 * len) return 65;
    std::fclose(f);
    std::printf("ctrl=%d first_direct=%d last_direct=%d\n", info[0], info[1], info[2]);
    return 0;
}
#endif
