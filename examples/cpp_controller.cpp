// C++ host program over include/vsmpc_adapter.hpp: one MPC instance driven exactly like the reference's
// closed-loop script drives `VariableSamplingMPC` (src/variable_sampling_mpc.py:68-71,106-131):
//   configure(...); per tick: update(pack); solveMPC(); read the getters; feed throttle / desired thrust /
//   desired thrust rate / joint reference back into the next tick's QPInput fields.
// The robot-side quantities of every tick (the pack) are read from a binary file written by the caller
// (tests/test_cpp_adapter.py); the outputs of every tick are written to another.
//
//   cpp_controller <in.bin> <out.bin>
//
// in.bin  (doubles): [alpha_len, traj_len, n_ticks, n_joints] | alphaGravity | positionCoM | velocityCoM | RPY |
//                    RPYDot (3 x traj_len each, sample-major) | jointPos[n_joints] | controlled[8] |
//                    configure pack[359] | n_ticks x pack[359]
// out.bin (doubles): n_ticks x ( throttle[4] thrust[4] thrustDot[4] jointsRef[n_joints] finalCoM[3] finalLinMom[3]
//                                finalRPY[3] finalAngMom[3] status )
#include <cstdio>
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
    if (argc < 3)
    {
        std::fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]);
        return 2;
    }
    std::vector<double> in;
    if (!read_all(argv[1], in) || in.size() < 4)
    {
        std::fprintf(stderr, "cannot read %s\n", argv[1]);
        return 2;
    }
    size_t o = 0;
    const int alphaLen = (int)in[o++], trajLen = (int)in[o++], nTicks = (int)in[o++], nJoints = (int)in[o++];
    vsmpc::Params p;
    auto take = [&](size_t n) { std::vector<double> r(in.begin() + o, in.begin() + o + n); o += n; return r; };
    p.alphaGravity = take(alphaLen);
    p.positionCoM = take(3 * (size_t)trajLen);
    p.velocityCoM = take(3 * (size_t)trajLen);
    p.RPY = take(3 * (size_t)trajLen);
    p.RPYDot = take(3 * (size_t)trajLen);
    std::vector<double> jointPos = take(nJoints);
    std::vector<int> controlled;
    for (double d : take(VSMPC_NJ))
        controlled.push_back((int)d);
    vsmpc::Pack pack;
    pack.set(0, in.data() + o, VSMPC_PACK_DOUBLES);
    o += VSMPC_PACK_DOUBLES;

    vsmpc::VariableSamplingMPC mpc;
    if (!mpc.configure(p, pack, jointPos, controlled))
        return 1;
    if (mpc.getNOptimizationVariables() != 588 || mpc.getNConstraints() != 512)
        std::fprintf(stderr, "note: nVar=%d nCon=%d\n", mpc.getNOptimizationVariables(), mpc.getNConstraints());

    std::vector<double> throttle(4), thrust(4), thrustDot(4), jointsRef(nJoints), v3(3), out;
    // the four QPInput fields the driver feeds back (src/variable_sampling_mpc.py:128-131); before the first
    // solve they hold what the caller put in the first pack
    for (int t = 0; t < nTicks; ++t)
    {
        pack.set(0, in.data() + o, VSMPC_PACK_DOUBLES);
        o += VSMPC_PACK_DOUBLES;
        if (t > 0)
        {
            pack.set(VSMPC_PK_THROTTLE_PREV, throttle.data(), 4);
            pack.set(VSMPC_PK_THRUST_DES, thrust.data(), 4);
            pack.set(VSMPC_PK_THRUST_DOT_DES, thrustDot.data(), 4);
            pack.setSelected(VSMPC_PK_Q_CMD, jointsRef.data(), controlled);
        }
        if (!mpc.update(pack) || !mpc.solveMPC())
            return 1;
        bool ok = mpc.getThrottleReference(throttle) && mpc.getThrustReference(thrust)
                  && mpc.getThrustDotReference(thrustDot) && mpc.getJointsReferencePosition(jointsRef);
        if (!ok)
            return 1;
        out.insert(out.end(), throttle.begin(), throttle.end());
        out.insert(out.end(), thrust.begin(), thrust.end());
        out.insert(out.end(), thrustDot.begin(), thrustDot.end());
        out.insert(out.end(), jointsRef.begin(), jointsRef.end());
        mpc.getFinalCoMPosition(v3);
        out.insert(out.end(), v3.begin(), v3.end());
        mpc.getFinalLinMom(v3);
        out.insert(out.end(), v3.begin(), v3.end());
        mpc.getFinalRPY(v3);
        out.insert(out.end(), v3.begin(), v3.end());
        mpc.getFinalAngMom(v3);
        out.insert(out.end(), v3.begin(), v3.end());
        out.push_back((double)mpc.getQPProblemStatus());
    }
    // size checks behave like the reference getters (variableSamplingMPC.cpp:138-151): wrong size -> false
    std::vector<double> wrong(5);
    if (mpc.getThrottleReference(wrong))
        return 3;
    FILE* f = std::fopen(argv[2], "wb");
    if (!f)
        return 2;
    std::fwrite(out.data(), sizeof(double), out.size(), f);
    std::fclose(f);
    std::printf("cpp_controller: %d ticks ok\n", nTicks);
    return 0;
}
