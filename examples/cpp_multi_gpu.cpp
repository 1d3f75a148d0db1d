// C++ host program over include/vsmpc_adapter.hpp: ONE process, ONE host thread, a batch of independent MPC instances
// packed instance by instance into structure-of-arrays buffers (vsmpc::PackBatch, north_star (a)) and sharded over several
// GPUs (vsmpc::MultiGpuMPC -> vsmpc_create_multi: contiguous ranges, one handle + stream per device, no inter-GPU traffic;
// the only gather is the per-device copy of the output rows).  The same batch also runs on one device through
// vsmpc::BatchedMPC; the program fails unless both give bit-identical rows on every tick.
//
//   cpp_multi_gpu <in.bin> <out.bin> <dev0,dev1,...>      e.g. 0,1 — or 0,0: two shards on one GPU
//
// in.bin  (doubles): [alpha_len, traj_len, n_ticks, n_joints, B] | alphaGravity | positionCoM | velocityCoM | RPY | RPYDot
//                    (3 x traj_len each, sample-major) | controlled[8] | B x jointPos[n_joints] |
//                    B x configure pack[359] (instance-major!) | n_ticks x B x pack[359]
// out.bin (doubles): n_ticks x B x ( row[54] status )
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "vsmpc_adapter.hpp"

static bool read_all(const char* path, std::vector<double>& v)
{
    FILE* f = std::fopen(path, "rb");
    if (!f)
        return false;
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    v.resize(n / sizeof(double));
    const size_t got = std::fread(v.data(), sizeof(double), v.size(), f);
    std::fclose(f);
    return got == v.size();
}

int main(int argc, char** argv)
{
    if (argc < 4)
    {
        std::fprintf(stderr, "usage: %s in.bin out.bin dev0,dev1,...\n", argv[0]);
        return 2;
    }
    std::vector<double> in;
    if (!read_all(argv[1], in) || in.size() < 5)
    {
        std::fprintf(stderr, "cannot read %s\n", argv[1]);
        return 2;
    }
    std::vector<int> devices;
    for (char* tok = std::strtok(argv[3], ","); tok; tok = std::strtok(nullptr, ","))
        devices.push_back(std::atoi(tok));
    size_t o = 0;
    const int alphaLen = (int)in[o++], trajLen = (int)in[o++], nTicks = (int)in[o++], nJoints = (int)in[o++], B = (int)in[o++];
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

    // every instance arrives as its own record (what a per-robot QPInput holds); PackBatch scatters it into SoA columns
    vsmpc::PackBatch batch(B);
    vsmpc::Pack pk;
    for (int i = 0; i < B; ++i)
        if (!batch.setJointPos(i, take(nJoints), controlled))
            return 1;
    for (int i = 0; i < B; ++i)
    {
        pk.set(0, in.data() + o, VSMPC_PACK_DOUBLES);
        o += VSMPC_PACK_DOUBLES;
        batch.set(i, pk);
    }
    std::vector<int> phase0(B);
    for (int i = 0; i < B; ++i)
        phase0[i] = i % 20; // staggered throttle-release phases: every tick has pinned and released instances

    vsmpc::MultiGpuMPC multi;
    vsmpc::BatchedMPC single;
    if (!multi.create(p, B, (int)devices.size(), devices) || !single.create(p, B, devices[0]))
        return 1;
    if (!multi.configure(batch, phase0.data()) || !single.configure(batch.pack(), batch.jointPosSel(), phase0.data()))
        return 1;
    int covered = 0;
    for (int g = 0; g < multi.nShards(); ++g)
    {
        int first = 0, count = 0;
        if (!multi.shard(g, first, count) || first != covered)
            return 4;
        covered += count;
    }
    if (covered != B)
        return 4;

    std::vector<double> rowsM((size_t)B * VSMPC_OUT_DOUBLES), rowsS(rowsM.size()), out;
    std::vector<int> statusM(B), statusS(B);
    for (int t = 0; t < nTicks; ++t)
    {
        for (int i = 0; i < B; ++i)
        {
            pk.set(0, in.data() + o, VSMPC_PACK_DOUBLES);
            o += VSMPC_PACK_DOUBLES;
            if (t > 0)
            { // feedback of the driver script (src/variable_sampling_mpc.py:128-131), per instance
                const double* r = rowsM.data() + (size_t)i * VSMPC_OUT_DOUBLES;
                pk.set(VSMPC_PK_THROTTLE_PREV, r + VSMPC_OUT_THROTTLE, 4);
                pk.set(VSMPC_PK_THRUST_DES, r + VSMPC_OUT_THRUST, 4);
                pk.set(VSMPC_PK_THRUST_DOT_DES, r + VSMPC_OUT_THRUST_DOT, 4);
                pk.set(VSMPC_PK_Q_CMD, r + VSMPC_OUT_JOINTS_REF, VSMPC_NJ);
            }
            batch.set(i, pk);
        }
        if (!multi.update(batch) || !multi.solveMPC() || !multi.getOutput(rowsM.data(), statusM.data()))
            return 1;
        if (!single.update(batch.pack()) || !single.solveMPC() || !single.getOutput(rowsS.data(), statusS.data()))
            return 1;
        if (std::memcmp(rowsM.data(), rowsS.data(), rowsM.size() * sizeof(double)) != 0
            || std::memcmp(statusM.data(), statusS.data(), B * sizeof(int)) != 0)
        {
            std::fprintf(stderr, "tick %d: sharded and single-device results differ\n", t);
            return 5;
        }
        for (int i = 0; i < B; ++i)
        {
            out.insert(out.end(), rowsM.begin() + (size_t)i * VSMPC_OUT_DOUBLES, rowsM.begin() + (size_t)(i + 1) * VSMPC_OUT_DOUBLES);
            out.push_back((double)statusM[i]);
        }
    }
    FILE* f = std::fopen(argv[2], "wb");
    if (!f)
        return 2;
    std::fwrite(out.data(), sizeof(double), out.size(), f);
    std::fclose(f);
    std::printf("cpp_multi_gpu: %d instances on %d shards, %d ticks, bit-identical to one device\n", B, (int)devices.size(), nTicks);
    return 0;
}
