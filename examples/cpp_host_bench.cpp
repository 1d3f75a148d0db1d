// C++ host layer end to end (north_star (a)): every tick the per-instance records (array of structures — what B robots' QPInput
// objects hold) are packed into the structure-of-arrays buffer (vsmpc::PackBatch::setMany, page-locked), uploaded, solved and the
// output rows read back, two ticks in flight (vsmpc_set_state + vsmpc_solve_async + vsmpc_get_output_async).  Prints one JSON line.
//   cpp_host_bench <input.bin> <steps> <warmup> <threads> [device]     input: the file tests/test_cpp_adapter.py writes for the
//                                                                       multi-GPU example (trajectories, joint positions, packs)
//   cpp_host_bench --selftest                                          PackBatch::setMany against PackBatch::set, no GPU needed
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "vsmpc_adapter.hpp"

using clk = std::chrono::steady_clock;
static double secs(clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); }

static int selftest()
{
    std::mt19937_64 g(7);
    std::uniform_real_distribution<double> u(-1.0, 1.0);
    vsmpc::PackThreads pool(4);
    for (int B : {1, 7, 8, 9, 250, 1024, 1031})
    {
        std::vector<vsmpc::Pack> recs(B);
        for (auto& r : recs)
            for (int k = 0; k < VSMPC_PACK_DOUBLES; ++k)
                r.v[k] = u(g);
        vsmpc::PackBatch a(B), b(B), c(B);
        for (int i = 0; i < B; ++i)
            a.set(i, recs[i]);
        b.setMany(0, recs.data(), B);
        const int off = B > 20 ? 5 : 0;                      // a range that starts off a 64-byte boundary, several threads
        for (int i = 0; i < off; ++i)
            c.set(i, recs[i]);
        c.setMany(off, recs.data() + off, B - off, &pool);
        const size_t n = (size_t)VSMPC_PACK_DOUBLES * B * sizeof(double);
        if (std::memcmp(a.pack(), b.pack(), n) != 0 || std::memcmp(a.pack(), c.pack(), n) != 0)
        {
            std::fprintf(stderr, "selftest: setMany differs from set at B = %d\n", B);
            return 1;
        }
    }
    std::printf("cpp_host_bench selftest ok\n");
    return 0;
}

static bool read_all(const char* path, std::vector<double>& v)
{
    FILE* f = std::fopen(path, "rb");
    if (!f)
        return false;
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    v.resize((size_t)n / sizeof(double));
    const size_t got = std::fread(v.data(), sizeof(double), v.size(), f);
    std::fclose(f);
    return got == v.size();
}

int main(int argc, char** argv)
{
    if (argc >= 2 && std::strcmp(argv[1], "--selftest") == 0)
        return selftest();
    if (argc < 5)
    {
        std::fprintf(stderr, "usage: cpp_host_bench <input.bin> <steps> <warmup> <threads> [device] | --selftest\n");
        return 2;
    }
    std::vector<double> in;
    if (!read_all(argv[1], in) || in.size() < 5)
        return 2;
    const int K = std::atoi(argv[2]), W = std::atoi(argv[3]), threads = std::atoi(argv[4]), device = argc > 5 ? std::atoi(argv[5]) : 0;
    size_t o = 0;
    const int alphaLen = (int)in[o++], trajLen = (int)in[o++], nSets = (int)in[o++], nJoints = (int)in[o++], B = (int)in[o++];
    vsmpc::Params p;
    auto take = [&](size_t n) { std::vector<double> r(in.begin() + o, in.begin() + o + n); o += n; return r; };
    p.alphaGravity = take(alphaLen);
    p.positionCoM = take(3 * (size_t)trajLen);
    p.velocityCoM = take(3 * (size_t)trajLen);
    p.RPY = take(3 * (size_t)trajLen);
    p.RPYDot = take(3 * (size_t)trajLen);
    std::vector<int> controlled;
    for (double d : take(VSMPC_NJ))
        controlled.push_back((int)d);
    vsmpc::PackBatch cfgBatch(B, true), bufA(B, true), bufB(B, true);
    vsmpc::PackBatch* buf[2] = {&bufA, &bufB};
    for (int i = 0; i < B; ++i)
        if (!cfgBatch.setJointPos(i, take(nJoints), controlled))
            return 1;
    // records: the nominal set (configure), then nSets perturbed sets the ticks cycle through
    std::vector<std::vector<vsmpc::Pack>> recs(1 + nSets, std::vector<vsmpc::Pack>(B));
    for (auto& set : recs)
        for (int i = 0; i < B; ++i)
        {
            set[i].set(0, in.data() + o, VSMPC_PACK_DOUBLES);
            o += VSMPC_PACK_DOUBLES;
        }
    vsmpc::PackThreads pool(threads);
    cfgBatch.setMany(0, recs[0].data(), B, &pool);
    std::vector<int> phase0(B);
    for (int i = 0; i < B; ++i)
        phase0[i] = i % 20;
    vsmpc::BatchedMPC mpc;
    if (!mpc.create(p, B, device) || !mpc.configure(cfgBatch.pack(), cfgBatch.jointPosSel(), phase0.data()))
        return 1;
    void* mem = nullptr;
    const size_t rowBytes = (size_t)B * VSMPC_OUT_DOUBLES * sizeof(double);
    if (vsmpc_host_alloc(2 * rowBytes + 2 * (size_t)B * sizeof(int), &mem) != VSMPC_OK)
        return 1;
    double* rows[2] = {static_cast<double*>(mem), static_cast<double*>(mem) + (size_t)B * VSMPC_OUT_DOUBLES};
    int* status[2] = {reinterpret_cast<int*>(static_cast<char*>(mem) + 2 * rowBytes),
                      reinterpret_cast<int*>(static_cast<char*>(mem) + 2 * rowBytes) + B};
    vsmpc_handle* h = mpc.handle();

    // pack alone: one instance at a time (strided columns) against the blocked scatter, one thread and `threads`
    auto time_pack = [&](int mode, int reps) {
        const auto t0 = clk::now();
        for (int r = 0; r < reps; ++r)
        {
            const std::vector<vsmpc::Pack>& set = recs[1 + r % nSets];
            if (mode == 0)
                for (int i = 0; i < B; ++i)
                    bufA.set(i, set[i]);
            else
                bufA.setMany(0, set.data(), B, mode == 1 ? nullptr : &pool);
        }
        return secs(t0, clk::now()) / reps;
    };
    time_pack(2, 3);
    const double tNaive = time_pack(0, 10), tBlocked1 = time_pack(1, 20), tBlockedT = time_pack(2, 20);

    double tPackSum = 0.0;
    auto loop = [&](int n, bool timed) {
        int prev = -1;
        for (int j = 0; j < n; ++j)
        {
            vsmpc::PackBatch& b = *buf[j & 1];       // the copy of tick j - 2 out of this buffer has landed: its output was waited for
            const auto p0 = clk::now();
            b.setMany(0, recs[1 + j % nSets].data(), B, &pool);
            if (timed)
                tPackSum += secs(p0, clk::now());
            int ticket = -1;
            if (vsmpc_set_state(h, b.pack()) != VSMPC_OK || vsmpc_solve_async(h) != VSMPC_OK
                || vsmpc_get_output_async(h, rows[j & 1], status[j & 1], &ticket) != VSMPC_OK)
                return false;
            if (prev >= 0 && vsmpc_wait_output(h, prev) != VSMPC_OK)
                return false;
            prev = ticket;
        }
        return prev < 0 || vsmpc_wait_output(h, prev) == VSMPC_OK;
    };
    if (!loop(W, false))
        return 1;
    const auto t0 = clk::now();
    if (!loop(K, true))
        return 1;
    const double t = secs(t0, clk::now());
    int solved = 0;
    for (int i = 0; i < B; ++i)
        solved += status[(K - 1) & 1][i] == 0;
    std::printf("{\"value\": %.1f, \"unit\": \"solves/s\", \"instances\": %d, \"steps\": %d, \"warmup\": %d, \"ms_per_step\": %.5f, "
                "\"host_threads\": %d, \"pack_ms_per_step\": %.5f, \"pack_alone_ms\": {\"one_instance_at_a_time\": %.5f, "
                "\"blocked_1_thread\": %.5f, \"blocked_threads\": %.5f}, \"pinned\": %s, \"solved_fraction_last_step\": %.4f, "
                "\"h2d_bytes_per_step\": %zu, \"d2h_bytes_per_step\": %zu}\n",
                (double)B * K / t, B, K, W, 1e3 * t / K, threads, 1e3 * tPackSum / K, 1e3 * tNaive, 1e3 * tBlocked1, 1e3 * tBlockedT,
                bufA.isPinned() ? "true" : "false", (double)solved / B, (size_t)VSMPC_PACK_DOUBLES * B * sizeof(double),
                rowBytes + (size_t)B * sizeof(int));
    vsmpc_host_free(mem);
    return 0;
}
