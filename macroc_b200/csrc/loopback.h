// loopback.h -- in-process communicator: N macroc_ctx of one process stand in for N ranks.
//
// The reference's one real invariant is decomposition independence (the same grid at
// -np 1,2,3,4,8, reference tests/CMakeLists.txt:21-28).  With NCCL that needs one GPU per rank;
// the loopback group lets N contexts -- each driven by its own host thread, on one device or
// several -- run the very same multi-rank code paths (z-slabs, x/y/PETSC_DECIDE boxes, the
// three-phase halo, the Gauss-point halo, the ghost-plane tiles of the symmetric operator):
//   * send/recv  = a device-to-device copy enqueued by the RECEIVER on its own stream, ordered
//                  after the sender's "data ready" event; the sender's stream then waits for the
//                  receiver's "copy done" event before it may touch the buffer again;
//   * all-reduce = every rank's partial goes to pinned host memory, is summed on the host in
//                  rank order (bit-reproducible) and copied back.
// Ranks meet at host barriers only (condition variable, with a timeout); no kernel ever waits
// for another kernel, so the contexts may share one GPU.
#pragma once

#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <vector>

#include <cuda_runtime.h>

namespace macroc {

struct LoopXfer {
    int peer;
    double *buf;
    size_t cnt;
};

struct LoopGroup {
    static constexpr const char *MAGIC = "MACROC-LOOPBACK";     // 15 chars + NUL = first 16 id bytes
    int n = 0;
    std::mutex mu;
    std::condition_variable cv;
    int arrived = 0, joined = 0, left = 0;
    uint64_t gen = 0;
    bool broken = false;
    struct Member {
        std::vector<LoopXfer> sends;          // published for the current exchange
        std::vector<size_t> taken;            // per peer: how many of my sends to it were consumed (scratch)
        cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
        int device = 0;
    };
    std::vector<Member> m;
    double *host_part = nullptr;              // pinned [n][4]
    int timeout_s = 120;

    explicit LoopGroup(int nranks) : n(nranks), m((size_t)nranks) {}

    // false: timed out or another member failed (the group is then unusable)
    bool barrier()
    {
        std::unique_lock<std::mutex> lk(mu);
        if (broken) return false;
        const uint64_t g = gen;
        if (++arrived == n) {
            arrived = 0; ++gen;
            cv.notify_all();
            return true;
        }
        cv.wait_for(lk, std::chrono::seconds(timeout_s), [&] { return gen != g || broken; });
        if (gen != g) return true;            // released (a member may already have left afterwards)
        broken = true;
        cv.notify_all();
        return false;
    }
    void fail()
    {
        std::lock_guard<std::mutex> lk(mu);
        broken = true;
        cv.notify_all();
    }
};

inline bool is_loopback_id(const void *id128)
{
    return id128 && memcmp(id128, LoopGroup::MAGIC, 16) == 0;
}
inline LoopGroup *loopback_group_of(const void *id128)
{
    LoopGroup *g = nullptr;
    memcpy(&g, (const unsigned char *)id128 + 16, sizeof(g));
    return g;
}

}  // namespace macroc
