// Row-sharded apps from C++ with no framework in the data path: two processes (fork), one GPU each, the
// vectors of the pull loops in a peer exchange (CUDA IPC, graphlily::Exchange) whose 64-byte handles travel
// over a socket pair; BFS / PageRank / SSSP pull against the oracle on the full matrix, repeated runs
// (the start-of-run barrier and the recorded sequence with the exchange inside).  Then the same apps over an
// NVSwitch multicast exchange the library creates itself (glb_xchg_mc_open / _bind): the multicast object's
// file descriptor travels from rank 0 to rank 1 as SCM_RIGHTS over the same socket pair.  Needs 2 GPUs.
#include <sys/socket.h>
#include <sys/wait.h>
#include <unistd.h>

#include "graphlily/app/bfs.h"
#include "graphlily/app/pagerank.h"
#include "graphlily/app/sssp.h"
#include "test_util.h"

using namespace graphlily;

static CSRMatrix<float> test_graph() {  // symmetric-ish skewed pattern, a few hubs, empty rows
    CSRMatrix<float> m = skewed_csr(6000, 77);
    for (auto &x : m.adj_data) x = 1.0f;
    return m;
}

static void swap_handles(int fd, const char *mine, char *all, int rank) {
    char peer[GLB_IPC_HANDLE_BYTES];
    if (write(fd, mine, GLB_IPC_HANDLE_BYTES) != GLB_IPC_HANDLE_BYTES || read(fd, peer, GLB_IPC_HANDLE_BYTES) != GLB_IPC_HANDLE_BYTES) {
        std::fprintf(stderr, "rank %d: handle exchange failed\n", rank);
        _exit(3);
    }
    std::memcpy(all + rank * GLB_IPC_HANDLE_BYTES, mine, GLB_IPC_HANDLE_BYTES);
    std::memcpy(all + (1 - rank) * GLB_IPC_HANDLE_BYTES, peer, GLB_IPC_HANDLE_BYTES);
}

static bool g_multicast = false;  // second pass of the worker: multicast exchange instead of CUDA IPC

static void sync_peers(int fd, int rank) {  // host barrier of the two processes
    char c = 'b';
    if (write(fd, &c, 1) != 1 || read(fd, &c, 1) != 1) {
        std::fprintf(stderr, "rank %d: barrier failed\n", rank);
        _exit(3);
    }
}

static void send_fd(int sock, int fd) {
    char byte = 'f', ctrl[CMSG_SPACE(sizeof(int))];
    std::memset(ctrl, 0, sizeof(ctrl));
    iovec io = {&byte, 1};
    msghdr msg = {};
    msg.msg_iov = &io;
    msg.msg_iovlen = 1;
    msg.msg_control = ctrl;
    msg.msg_controllen = sizeof(ctrl);
    cmsghdr *c = CMSG_FIRSTHDR(&msg);
    c->cmsg_level = SOL_SOCKET;
    c->cmsg_type = SCM_RIGHTS;
    c->cmsg_len = CMSG_LEN(sizeof(int));
    std::memcpy(CMSG_DATA(c), &fd, sizeof(int));
    if (sendmsg(sock, &msg, 0) != 1) _exit(3);
}

static int recv_fd(int sock) {
    char byte = 0, ctrl[CMSG_SPACE(sizeof(int))];
    std::memset(ctrl, 0, sizeof(ctrl));
    iovec io = {&byte, 1};
    msghdr msg = {};
    msg.msg_iov = &io;
    msg.msg_iovlen = 1;
    msg.msg_control = ctrl;
    msg.msg_controllen = sizeof(ctrl);
    if (recvmsg(sock, &msg, 0) != 1) _exit(3);
    cmsghdr *c = CMSG_FIRSTHDR(&msg);
    if (!c || c->cmsg_type != SCM_RIGHTS) _exit(3);
    int fd = -1;
    std::memcpy(&fd, CMSG_DATA(c), sizeof(int));
    return fd;
}

static std::unique_ptr<Exchange> open_multicast(std::shared_ptr<Runtime> rt, uint32_t n, int rank, int sock) {
    std::unique_ptr<Exchange> xc;
    int fd = -1;
    if (rank == 0) {
        xc.reset(new Exchange(rt, n, 3, 0, 2, -1, &fd));
        send_fd(sock, fd);
    } else {
        fd = recv_fd(sock);
        xc.reset(new Exchange(rt, n, 3, 1, 2, fd, nullptr));
    }
    close(fd);
    sync_peers(sock, rank);  // both devices are in the multicast object
    xc->bind();
    sync_peers(sock, rank);
    EXPECT_TRUE(xc->has_multicast());
    return xc;
}

template <typename App>
static std::unique_ptr<Exchange> shard(App &app, std::shared_ptr<Runtime> rt, uint32_t n, int rank, int fd) {
    if (g_multicast) {
        std::unique_ptr<Exchange> mc = open_multicast(rt, n, rank, fd);
        app.set_sharding(rank, 2, mc.get());
        app.send_matrix_host_to_device();
        return mc;
    }
    std::unique_ptr<Exchange> xc(new Exchange(rt, n, 3));
    char mine[GLB_IPC_HANDLE_BYTES], all[2 * GLB_IPC_HANDLE_BYTES];
    xc->export_handle(mine);
    swap_handles(fd, mine, all, rank);
    xc->connect(rank, 2, all);
    app.set_sharding(rank, 2, xc.get());
    app.send_matrix_host_to_device();
    return xc;
}

static void run_apps(std::shared_ptr<Runtime> rt, int rank, int fd);

static int worker(int rank, int fd) {
    std::shared_ptr<Runtime> rt = std::make_shared<Runtime>(rank);
    run_apps(rt, rank, fd);
    if (Exchange::multicast_supported(rt)) {
        g_multicast = true;
        run_apps(rt, rank, fd);
        if (rank == 0) std::printf("multicast exchange (glb_xchg_mc_open / _bind): apps re-run over it\n");
    } else if (rank == 0) {
        std::printf("no multicast support on this device: second pass skipped\n");
    }
    return mini_test::failures();
}

static void run_apps(std::shared_ptr<Runtime> rt, int rank, int fd) {
    auto g = test_graph();
    {
        CSRMatrix<float> m = g;
        io::util_round_csr_matrix_dim(m, 128, 128);
        for (auto &x : m.adj_data) x = 1;
        app::BFS bfs(16, 1024, 512, 256);
        bfs.set_runtime(rt);
        bfs.load_and_format_matrix(g, true);
        auto xc = shard(bfs, rt, m.num_rows, rank, fd);
        for (uint32_t src : {0u, 7u, 0u}) {
            dense_t ref(m.num_rows);
            oracle_bfs(m.num_rows, m.adj_indptr.data(), m.adj_indices.data(), m.adj_data.data(), src, 6, ref.data());
            verify(ref, bfs.pull(src, 6), true);
            verify(ref, bfs.push(src, 6), true);   // CSC row shards, the frontier exchanged as a dense vector
            for (float thr : {0.002f, 0.2f, 1.1f}) verify(ref, bfs.pull_push(src, 6, thr), true);
        }
        rt->finish();
    }
    {
        // short rows: the skewed graph's 5000-long rows put the REFERENCE's sequential fp32 sum 1.3e-5 off
        // (DESIGN.md section 4); the 1e-5 bar is checked where the summation order cannot matter
        auto gu = uniform_csr(6000, 12, 5, 1.0f);
        CSRMatrix<float> m = gu;
        io::util_round_csr_matrix_dim(m, 128, 128);
        io::util_normalize_csr_matrix_by_outdegree(m);
        for (auto &x : m.adj_data) x = x * 0.9f;
        app::PageRank pr(16, 1024, 256);
        pr.set_runtime(rt);
        pr.load_and_format_matrix(gu, 0.9f, true);
        auto xc = shard(pr, rt, m.num_rows, rank, fd);
        for (uint32_t iters : {5u, 6u, 5u}) {
            dense_t ref(m.num_rows);
            oracle_pagerank(m.num_rows, m.adj_indptr.data(), m.adj_indices.data(), m.adj_data.data(), 0.9f, iters, ref.data());
            verify(ref, pr.pull(0.9f, iters), false);
        }
        rt->finish();
    }
    {
        CSRMatrix<float> m = g;
        app::detail::sssp_preprocess(m);
        io::util_round_csr_matrix_dim(m, 128, 128);
        app::SSSP sssp(16, 1024, 512, 256);
        sssp.set_runtime(rt);
        sssp.load_and_format_matrix(g, true);
        auto xc = shard(sssp, rt, m.num_rows, rank, fd);
        for (uint32_t src : {0u, 11u}) {
            dense_t ref(m.num_rows);
            oracle_sssp(m.num_rows, m.adj_indptr.data(), m.adj_indices.data(), m.adj_data.data(), src, 5, TropicalSemiring.zero,
                        ref.data());
            verify(ref, sssp.pull(src, 5), true);
            verify(ref, sssp.push(src, 5), true);
            for (float thr : {0.002f, 0.2f, 1.1f}) verify(ref, sssp.pull_push(src, 5, thr), true);
        }
        rt->finish();
    }
}

int main() {
    int n_dev = 0;
    // the parent never touches CUDA (a forked child could not use it afterwards): ask a child for the device count
    int fds[2], probe[2];
    if (pipe(probe) != 0) return 2;
    pid_t p = fork();
    if (p == 0) {
        int n = 0;
        glb_device_count(&n);
        if (write(probe[1], &n, sizeof(n)) != sizeof(n)) _exit(1);
        _exit(0);
    }
    if (read(probe[0], &n_dev, sizeof(n_dev)) != sizeof(n_dev)) n_dev = 0;
    waitpid(p, nullptr, 0);
    if (n_dev < 2) {
        std::printf("SKIPPED: needs 2 GPUs (found %d)\n0 test(s) failed\n", n_dev);
        return 0;
    }
    if (socketpair(AF_UNIX, SOCK_STREAM, 0, fds) != 0) return 2;
    pid_t kids[2];
    for (int r = 0; r < 2; r++) {
        kids[r] = fork();
        if (kids[r] == 0) {
            close(fds[1 - r]);
            int bad = 1;
            try { bad = worker(r, fds[r]); } catch (std::exception &e) { std::fprintf(stderr, "rank %d: %s\n", r, e.what()); } catch (mini_test::Abort &) {}
            std::fflush(stdout);
            _exit(bad ? 1 : 0);
        }
    }
    int failed = 0;
    for (int r = 0; r < 2; r++) {
        int st = 0;
        waitpid(kids[r], &st, 0);
        const bool ok = WIFEXITED(st) && WEXITSTATUS(st) == 0;
        std::printf("[%s] Sharded.Rank%d (BFS / SSSP pull, push, pull_push and PageRank over a CUDA-IPC peer exchange, then over a library-made multicast exchange; 2 processes)\n", ok ? "  OK  " : "FAILED", r);
        failed += !ok;
    }
    std::printf("%d test(s) failed\n", failed);
    return failed;
}
