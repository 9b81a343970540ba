// rendezvous.cu -- host-only control plane for sharded registers: all-gather of small blobs and barriers between the
// processes of ONE node through files in a shared directory (/dev/shm by default).
//
// The data plane needs none of this (exchanges are peer loads / stores by our own kernels, dist.cu); what the processes must
// trade on the host is 256 bytes of CUDA IPC handles each at start-up, a barrier around timed regions and a few scalars.  A file
// rendezvous keeps that out of any framework: the Python mirror (spinoza_b200/distributed.py), a C++ or a Rust caller use the
// same four functions, and nothing on the product path imports torch.
//
// Protocol.  Operation number `seq` of rank r publishes <dir>/<seq>.<r> (written under a temporary name, then renamed, so a
// reader never sees a partial file) and polls for the files of the other ranks.  A rank removes its file of operation seq - 2
// when it starts operation seq: by then every rank has finished reading operation seq - 2 (it could not have published
// seq - 1 otherwise, and this rank has seen all files of seq - 1).
#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include "engine.h"

struct spz_rdv {
    std::string dir;
    int rank = 0, world = 1;
    long long seq = 0;
    int timeout_ms = 600000;
};

using namespace spz;

static std::string rdv_path(const spz_rdv *r, long long seq, int rank) {
    return r->dir + "/" + std::to_string(seq) + "." + std::to_string(rank);
}

extern "C" {

int spz_rdv_open(const char *dir, int rank, int world, spz_rdv **out) {
    if (!out || world < 1 || rank < 0 || rank >= world) { set_error("spz_rdv_open: bad rank / world"); return SPZ_ERR_INVALID_ARG; }
    *out = nullptr;
    spz_rdv *r = new spz_rdv();
    r->rank = rank; r->world = world;
    if (dir && dir[0]) {
        r->dir = dir;
    } else {
        // one directory per launch: the launcher's port and its process id are the same for every rank of a torchrun /
        // mpirun-style launch on one node and differ between launches
        const char *port = std::getenv("MASTER_PORT");
        r->dir = std::string("/dev/shm/spz_rdv_") + (port ? port : "0") + "_" + std::to_string((long long)getppid()) + "_" +
                 std::to_string((long long)getuid());
    }
    if (const char *t = std::getenv("SPZ_RDV_TIMEOUT_MS")) { const int v = std::atoi(t); if (v > 0) r->timeout_ms = v; }
    if (mkdir(r->dir.c_str(), 0700) != 0 && errno != EEXIST) {
        set_error("spz_rdv_open: cannot create %s: %s", r->dir.c_str(), std::strerror(errno));
        delete r;
        return SPZ_ERR_COMM;
    }
    *out = r;
    return SPZ_OK;
}

// every rank contributes `bytes` bytes; `all` receives world * bytes in rank order
int spz_rdv_allgather(spz_rdv *r, const void *mine, int64_t bytes, void *all) {
    if (!r || bytes < 0 || (bytes && (!mine || !all))) { set_error("spz_rdv_allgather: bad arguments"); return SPZ_ERR_INVALID_ARG; }
    const long long seq = r->seq++;
    if (seq >= 2) unlink(rdv_path(r, seq - 2, r->rank).c_str());
    const std::string final_name = rdv_path(r, seq, r->rank), tmp = final_name + ".tmp";
    FILE *f = std::fopen(tmp.c_str(), "wb");
    if (!f || (bytes && std::fwrite(mine, 1, (size_t)bytes, f) != (size_t)bytes) || std::fclose(f) != 0 ||
        std::rename(tmp.c_str(), final_name.c_str()) != 0) {
        set_error("spz_rdv: cannot publish %s: %s", final_name.c_str(), std::strerror(errno));
        return SPZ_ERR_COMM;
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (int p = 0; p < r->world; ++p) {
        char *dst = static_cast<char *>(all) + (size_t)p * (size_t)bytes;
        if (p == r->rank) { if (bytes) std::memcpy(dst, mine, (size_t)bytes); continue; }
        const std::string name = rdv_path(r, seq, p);
        int spins = 0;
        for (;;) {
            FILE *g = std::fopen(name.c_str(), "rb");
            if (g) {
                const size_t got = bytes ? std::fread(dst, 1, (size_t)bytes, g) : 0;
                std::fclose(g);
                if (got == (size_t)bytes) break;
                set_error("spz_rdv: rank %d published %zu bytes, expected %lld", p, got, (long long)bytes);
                return SPZ_ERR_COMM;
            }
            if (std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count() > r->timeout_ms) {
                set_error("spz_rdv: rank %d did not reach operation %lld within %d ms", p, seq, r->timeout_ms);
                return SPZ_ERR_COMM;
            }
            if (++spins < 200) std::this_thread::yield();
            else std::this_thread::sleep_for(std::chrono::microseconds(spins < 2000 ? 50 : 1000));
        }
    }
    return SPZ_OK;
}

int spz_rdv_barrier(spz_rdv *r) {
    char mine = 0, all[64];
    std::vector<char> big;
    char *dst = all;
    if (r && r->world > 64) { big.resize((size_t)r->world); dst = big.data(); }
    return spz_rdv_allgather(r, &mine, 1, dst);
}

int spz_rdv_close(spz_rdv *r) {
    if (!r) return SPZ_OK;
    // A rank may delete a file only when it knows that every rank has read it, and it learns that from the NEXT operation's
    // files.  (The first version unlinked its own file of the final barrier as soon as the barrier returned: a rank that
    // published a moment later found the file gone and polled until the time-out -- a 600 s hang at exit, seen on 2 of 7 two-GPU
    // runs.)  So closing is: a barrier S; then every rank but 0 publishes one more file, S + 1, and leaves without waiting;
    // rank 0 waits for those files -- whoever published S + 1 has passed S, i.e. has read everything -- and removes all that
    // is left, the directory included.
    if (r->world > 1 && spz_rdv_barrier(r) == SPZ_OK) {
        const long long fin = r->seq; // S + 1
        if (r->rank != 0) {
            const std::string name = rdv_path(r, fin, r->rank), tmp = name + ".tmp";
            if (FILE *f = std::fopen(tmp.c_str(), "wb")) { std::fclose(f); std::rename(tmp.c_str(), name.c_str()); }
        } else {
            const auto t0 = std::chrono::steady_clock::now();
            bool all_left = true;
            for (int p = 1; p < r->world && all_left; ++p) {
                const std::string name = rdv_path(r, fin, p);
                while (access(name.c_str(), F_OK) != 0) {
                    if (std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count() > 10000) { all_left = false; break; }
                    std::this_thread::sleep_for(std::chrono::microseconds(200));
                }
            }
            if (all_left) {
                for (long long q = fin - 3 < 0 ? 0 : fin - 3; q <= fin; ++q)
                    for (int p = 0; p < r->world; ++p) unlink(rdv_path(r, q, p).c_str());
                rmdir(r->dir.c_str());
            }
        }
    }
    delete r;
    return SPZ_OK;
}

// export + all-gather + connect: everything a sharded register needs before its first gate
int spz_dist_connect_rdv(spz_state *st, spz_rdv *r) {
    if (!st || !r) { set_error("spz_dist_connect_rdv: null argument"); return SPZ_ERR_INVALID_ARG; }
    std::vector<char> mine(SPZ_IPC_BLOB_BYTES), all((size_t)SPZ_IPC_BLOB_BYTES * (size_t)r->world);
    SPZ_TRY(spz_dist_export(st, mine.data()));
    SPZ_TRY(spz_rdv_allgather(r, mine.data(), SPZ_IPC_BLOB_BYTES, all.data()));
    SPZ_TRY(spz_dist_connect(st, all.data()));
    return spz_rdv_barrier(r);
}

} // extern "C"
