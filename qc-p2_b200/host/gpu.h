// gpu.h -- process-wide qcp2 context for the host mirror (host pointer mode) and the optional
// multi-GPU worker pool of `P2 --gpus N`: one process per GPU (rank r drives device r), forked
// before any CUDA or OpenMP state exists; rank 0 runs the reference's program flow, the workers
// wait on a pipe and join the collective (T) evaluation (qcp2_ccsd_t_mgpu) when it is reached.
#pragma once
#include <sys/types.h>
#include <sys/wait.h>
#include <signal.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/qcp2_b200.h"

struct MgpuState {
    int world = 1, rank = 0;
    std::vector<int> to_worker;// rank 0: write ends of the worker pipes
    std::vector<pid_t> pids;
};
inline MgpuState &mgpu() {
    static MgpuState s;
    return s;
}
struct MgpuCommand {
    int cmd;// 0 = exit, 1 = join qcp2_ccsd_t_mgpu, 2 = join qcp2_post_hf_mgpu
    int o, v;
    int n, uhf, stages, maxiter;// cmd 2
    double conv;
    char id[QCP2_COMM_ID_BYTES];
};

inline qcp2_ctx *&gpu_slot() {
    static qcp2_ctx *c = nullptr;
    return c;
}
inline qcp2_ctx *gpu() {
    qcp2_ctx *&c = gpu_slot();
    if (!c) {
        const int rc = qcp2_create(mgpu().rank, &c);
        if (rc != 0) throw std::runtime_error("qcp2_b200: no usable B200 device (status " + std::to_string(rc) + "); there is no CPU fallback");
    }
    return c;
}
inline void gpu_check(int rc) {
    if (rc != 0) throw std::runtime_error(std::string("qcp2_b200: ") + qcp2_last_error(gpu()));
}
inline void gpu_release() {
    qcp2_ctx *&c = gpu_slot();
    if (c) qcp2_destroy(c), c = nullptr;
}

// worker side: serve commands until told to exit; never returns
[[noreturn]] inline void mgpu_worker_loop(int fd) {
    int status = 0;
    for (;;) {
        MgpuCommand m;
        size_t got = 0;
        while (got < sizeof(m)) {
            const ssize_t r = read(fd, reinterpret_cast<char *>(&m) + got, sizeof(m) - got);
            if (r <= 0) _exit(status);// parent went away
            got += (size_t) r;
        }
        if (m.cmd == 0) break;
        try {
            double e = 0.0;
            gpu_check(qcp2_comm_init(gpu(), m.id, mgpu().world, mgpu().rank));
            if (m.cmd == 1) {
                gpu_check(qcp2_ccsd_t_mgpu(gpu(), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, m.o, m.v, QCP2_T_AUTO, 0,
                                           nullptr, &e));
            } else {
                double e2 = 0.0, ecc = 0.0;
                int iters = 0;
                gpu_check(qcp2_post_hf_mgpu(gpu(), nullptr, nullptr, nullptr, nullptr, nullptr, m.n, m.o, m.v, m.uhf, m.stages, m.maxiter,
                                            m.conv, 0, &e2, &ecc, &e, &iters, nullptr, nullptr, nullptr));
            }
            gpu_check(qcp2_comm_destroy(gpu()));
        } catch (const std::exception &ex) {
            fprintf(stderr, "P2 worker %d: %s\n", mgpu().rank, ex.what());
            status = 1;
        }
    }
    gpu_release();
    _exit(status);
}

// fork world-1 workers; returns in rank 0 only
inline void mgpu_spawn(int world) {
    if (world <= 1) return;
    {// count the GPUs in a throw-away child, so that this process still forks without a CUDA context
        const pid_t probe = fork();
        if (probe < 0) throw std::runtime_error("P2 --gpus: fork() failed");
        if (probe == 0) _exit(std::min(qcp2_device_count(), 255));
        int st = 0;
        waitpid(probe, &st, 0);
        const int ndev = WIFEXITED(st) ? WEXITSTATUS(st) : 0;
        if (ndev < world) throw std::runtime_error("P2 --gpus " + std::to_string(world) + ": only " + std::to_string(ndev) + " GPU(s) visible");
    }
    MgpuState &s = mgpu();
    s.world = world;
    for (int r = 1; r < world; ++r) {
        int fds[2];
        if (pipe(fds) != 0) throw std::runtime_error("P2 --gpus: pipe() failed");
        const pid_t pid = fork();
        if (pid < 0) throw std::runtime_error("P2 --gpus: fork() failed");
        if (pid == 0) {
            close(fds[1]);
            for (int w : s.to_worker) close(w);
            s.rank = r;
            mgpu_worker_loop(fds[0]);
        }
        close(fds[0]);
        s.to_worker.push_back(fds[1]);
        s.pids.push_back(pid);
    }
}
inline void mgpu_send(const MgpuCommand &m) {
    for (int fd : mgpu().to_worker)
        if (write(fd, &m, sizeof(m)) != (ssize_t) sizeof(m)) throw std::runtime_error("P2 --gpus: worker pipe closed");
}
// rank 0: dismiss the workers and collect their exit codes
inline int mgpu_shutdown() {
    MgpuState &s = mgpu();
    int bad = 0;
    if (s.rank == 0 && !s.pids.empty()) {
        MgpuCommand m;
        std::memset(&m, 0, sizeof(m));
        for (int fd : s.to_worker) {
            if (write(fd, &m, sizeof(m)) != (ssize_t) sizeof(m)) bad = 1;
            close(fd);
        }
        // a worker that is still inside a collective because rank 0 failed mid-way never reads the dismissal:
        // wait a bounded time (QCP2_WORKER_TIMEOUT_S, default 60 s), then kill it -- `P2 --gpus N` must exit
        const char *tenv = getenv("QCP2_WORKER_TIMEOUT_S");
        const double limit = tenv ? atof(tenv) : 60.0;
        for (pid_t p : s.pids) {
            int st = 0;
            double waited = 0.0;
            pid_t r = 0;
            while ((r = waitpid(p, &st, WNOHANG)) == 0 && waited < limit) {
                usleep(20000);
                waited += 0.02;
            }
            if (r == 0) {
                fprintf(stderr, "P2: worker %d did not exit within %.0f s, killing it\n", (int) p, limit);
                kill(p, SIGKILL);
                waitpid(p, &st, 0);
                bad = 1;
            } else if (r < 0 || !WIFEXITED(st) || WEXITSTATUS(st) != 0) {
                bad = 1;
            }
        }
        s.to_worker.clear(), s.pids.clear();
    }
    return bad;
}
