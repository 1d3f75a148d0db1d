// vsmpc_reference_glue.hpp — the reference-side binding of libvsmpc: a class with EXACTLY the reference's
// configure / update / solveMPC / getter signatures that a maintainer drops into the reference tree.
//
// Compile it INSIDE the reference (it includes the reference's own QPInput.h / Robot.h / JetModel.h and the libraries they
// use: Eigen, iDynTree, BipedalLocomotion's parameter handler, YARP's ResourceFinder, matio) and link libvsmpc.so:
//
//     #include <vsmpc_reference_glue.hpp>
//     vsmpc::VariableSamplingMPCOnGpu mpc;                      // instead of ::VariableSamplingMPC
//     mpc.configure(parametersHandler, qpInput);                // IMPCProblem::configure        IMPCProblem.h:35
//     mpc.update(qpInput); mpc.solveMPC();                      // IMPCProblem::update :44, VariableSamplingMPC::solveMPC
//     mpc.getThrustReference(thrust); ...                       // variableSamplingMPC.h:18-40
//
// What it does per call (nothing of the MPC is computed on the host):
//   configure : reads the parameters the reference's costs / constraints read (same names, same "not found" errors:
//               variableSamplingMPC.cpp:12-36, costsVSMPC.cpp:29-69,326-347,445-455,516-531, constraintsVSMPC.cpp:25-40,
//               171-176,295-317, systemDynamicsVSMPC.cpp:18-30,240-272,362-372), loads the two trajectory files exactly like
//               TrajectoryManager::loadTrajectoryFromFile (TrajectoryManager.cpp:67-140: ResourceFinder + matio, raw samples —
//               the library resamples), resolves the controlled joints by name (variableSamplingMPC.cpp:25-37), packs the
//               robot state (fillPack) and calls vsmpc_create + vsmpc_configure;
//   update    : fillPack(QPInput) -> vsmpc_set_state, then writes the four QPInput fields the reference's update() writes
//               (costsVSMPC.cpp:155-160 setPosCoMReference / setRPYReference / setMomentumReference,
//               systemDynamicsVSMPC.cpp:310 setAlphaGravity) from vsmpc_get_references;
//   solveMPC  : vsmpc_solve + vsmpc_get_output; returns true also for a non-solved status (variableSamplingMPC.cpp:111).
//
// Only element access, rows()/cols()/size() are used on Eigen / iDynTree objects, so the header also compiles against the
// stand-in headers of oracle/ref_stubs/ — which is how tests/test_reference_glue.py builds and runs it here, against the
// reference's own QPInput.cpp and side by side with the reference's own VariableSamplingMPC on the same QPInput object.
#ifndef VSMPC_REFERENCE_GLUE_HPP
#define VSMPC_REFERENCE_GLUE_HPP

#include <cmath>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <BipedalLocomotion/ParametersHandler/YarpImplementation.h>
#include <Eigen/Dense>
#include <iDynTree/EigenHelpers.h>
#include <matio.h>
#include <yarp/os/LogStream.h>
#include <yarp/os/ResourceFinder.h>

#include <QPInput.h> // the reference's: Robot.h, JetModel.h

#include "vsmpc_adapter.hpp"

namespace vsmpc
{

using ParametersHandlerWeakPtr = std::weak_ptr<BipedalLocomotion::ParametersHandler::IParametersHandler>;

namespace glue
{
// ---- jet model constants (private members of the reference's JetModel, UT/include/JetModel.h:93-94) recovered through its
// public functions; the literals of vsmpc::Params (= JetModel.cpp:13-26) are kept when the instance agrees with them to
// 1e-12, so that the default model is bit-identical
inline void jetConstants(JetModel& jm, double (&coeff)[13], double (&norm)[4])
{
    double c[13], n[4];
    c[0] = jm.compute_f(0.0, 0.0);
    c[1] = jm.compute_df_dT(0.0, 0.0);
    c[2] = jm.compute_df_dTdot(0.0, 0.0);
    c[3] = jm.compute_df_dT(0.0, 1.0) - c[1];
    c[4] = 0.5 * (jm.compute_df_dT(1.0, 0.0) - c[1]);
    c[5] = 0.5 * (jm.compute_df_dTdot(0.0, 1.0) - c[2]);
    c[6] = jm.compute_g(0.0, 0.0);
    c[7] = jm.compute_dg_dT(0.0, 0.0);
    c[8] = jm.compute_dg_dTdot(0.0, 0.0);
    c[9] = jm.compute_dg_dT(0.0, 1.0) - c[7];
    c[10] = 0.5 * (jm.compute_dg_dT(1.0, 0.0) - c[7]);
    c[11] = 0.5 * (jm.compute_dg_dTdot(0.0, 1.0) - c[8]);
    c[12] = jm.compute_v(1.0) - 1.0;
    n[0] = jm.destandardizeThrust_u2T(0.0);
    n[1] = jm.getThrustStandardDeviation_u2T();
    const double s0 = jm.standardizeThrottle_u2T(0.0), s1 = jm.standardizeThrottle_u2T(1.0);
    n[3] = 1.0 / (s1 - s0);
    n[2] = -s0 * n[3];
    bool same = true;
    for (int i = 0; i < 13; ++i)
        same = same && std::fabs(c[i] - coeff[i]) <= 1e-12 * std::fmax(1.0, std::fabs(coeff[i]));
    for (int i = 0; i < 4; ++i)
        same = same && std::fabs(n[i] - norm[i]) <= 1e-12 * std::fmax(1.0, std::fabs(norm[i]));
    if (!same)
    {
        for (int i = 0; i < 13; ++i)
            coeff[i] = c[i];
        for (int i = 0; i < 4; ++i)
            norm[i] = n[i];
    }
}

// TrajectoryManager::loadTrajectoryFromFile (TrajectoryManager.cpp:67-140) without the resampling: every variable of the
// file as raw samples (MATLAB column-major dim x samples = sample-major [dim * s + axis]) + the file's "fps"
inline bool loadTrajectoryFile(const std::string& filename, std::map<std::string, std::vector<double>>& vars,
                               std::map<std::string, int>& dims, int& fps)
{
    yarp::os::ResourceFinder rf;
    const std::string path = rf.findFileByName(filename);
    mat_t* mat = Mat_Open(path.c_str(), MAT_ACC_RDONLY);
    if (mat == nullptr)
    {
        yError() << "Error opening file " << filename;
        return false;
    }
    matvar_t* fpsVar = Mat_VarRead(mat, "fps");
    if (fpsVar == nullptr)
    {
        yError() << "Error reading fps";
        return false;
    }
    fps = static_cast<int>(*static_cast<double*>(fpsVar->data));
    for (matvar_t* info = Mat_VarReadNextInfo(mat); info != nullptr; info = Mat_VarReadNextInfo(mat))
    {
        matvar_t* v = Mat_VarRead(mat, info->name);
        if (v == nullptr)
        {
            yError() << "Error reading " << info->name << " variable";
            return false;
        }
        const double* d = static_cast<const double*>(v->data);
        vars[info->name] = std::vector<double>(d, d + v->dims[0] * v->dims[1]);
        dims[info->name] = static_cast<int>(v->dims[0]);
    }
    return true;
}

inline bool trajectoryFileOfGroup(const std::shared_ptr<BipedalLocomotion::ParametersHandler::IParametersHandler>& ptr,
                                  const char* group, std::string& file)
{
    auto g = ptr->getGroup(group).lock();
    if (g == nullptr)
    {
        yError() << "Group [TRAJECTORY_MANAGER] not found in the config file."; // (the reference's message for both groups)
        return false;
    }
    if (!g->getParameter("trajectoryFile", file))
    {
        yError() << "Unable to read the name of the trajectory.";
        return false;
    }
    return true;
}

template <class T>
inline bool need(const std::shared_ptr<BipedalLocomotion::ParametersHandler::IParametersHandler>& ptr, const char* name, T& v)
{
    if (!ptr->getParameter(name, v))
    {
        yError() << "Parameter '" << name << "' not found in the config file.";
        return false;
    }
    return true;
}

inline bool weight3(const std::shared_ptr<BipedalLocomotion::ParametersHandler::IParametersHandler>& ptr, const char* name,
                    double (&w)[3])
{
    std::vector<double> v;
    if (!ptr->getParameter(name, v) || v.size() != 3)
    { // getParameterAndCheckSize, costsVSMPC.cpp:9-23
        yError() << "Parameter '" << name << "' not found in the config file or of the wrong size.";
        return false;
    }
    for (int a = 0; a < 3; ++a)
        w[a] = v[a];
    return true;
}
} // namespace glue

// the parameters of group VS_MPC_CONFIG the reference's classes read + the two trajectory files they name
inline bool paramsFromHandler(ParametersHandlerWeakPtr parametersHandler, JetModel* jetModel, Params& p,
                              std::vector<std::string>& controlledJoints)
{
    auto ptr = parametersHandler.lock();
    if (ptr == nullptr)
    {
        yError() << "vsmpc::paramsFromHandler: the parameter handler is expired";
        return false;
    }
    using glue::need;
    bool ok = need(ptr, "controlledJoints", controlledJoints) && need(ptr, "nIter", p.nIter) && need(ptr, "nIterSmall", p.nIterSmall)
              && need(ptr, "controlHorizon", p.controlHorizon) && need(ptr, "periodMPC", p.periodMPC)
              && need(ptr, "periodMPCLargeSteps", p.periodMPCLargeSteps) && need(ptr, "periodMPCSmallSteps", p.periodMPCSmallSteps)
              && need(ptr, "useJetDynamic", p.useJetDynamic) && need(ptr, "useEstimatedThrust", p.useEstimatedThrust)
              && need(ptr, "jointsLambdaOption", p.jointsLambdaOption) && need(ptr, "weightThrottle", p.weightThrottle)
              && need(ptr, "weightInitialThrottle", p.weightInitialThrottle)
              && need(ptr, "weightRegularizationJointPos", p.weightRegularizationJointPos) && need(ptr, "throttleMax", p.throttleMax)
              && need(ptr, "throttleMin", p.throttleMin);
    ok = ok && glue::weight3(ptr, "weightCoMPos", p.weightCoMPos) && glue::weight3(ptr, "weightCoMPosError", p.weightCoMPosError)
         && glue::weight3(ptr, "weightLinMom", p.weightLinMom) && glue::weight3(ptr, "weightRPY", p.weightRPY)
         && glue::weight3(ptr, "weightRPYError", p.weightRPYError) && glue::weight3(ptr, "weightAngMom", p.weightAngMom);
    if (!ok)
        return false;
    if (p.jointsLambdaOption != "unfiltered" && p.jointsLambdaOption != "constant")
    { // systemDynamicsVSMPC.cpp:30-46
        yError() << "Parameter 'jointsLambdaOption' should be 'unfiltered' or 'constant'.";
        return false;
    }
    std::vector<double> wdj;
    if (!need(ptr, "weightDeltaJoint", wdj))
        return false;
    if (wdj.size() != VSMPC_NJ)
    { // costsVSMPC.cpp:352-357
        yError() << "The size of the vector containing the weights for the joint deltas is not correct.";
        return false;
    }
    for (int a = 0; a < VSMPC_NJ; ++a)
        p.weightDeltaJoint[a] = wdj[a];
    if (jetModel != nullptr)
        glue::jetConstants(*jetModel, p.jetCoeff, p.jetNorm);
    // TRAJECTORY_MANAGER -> alphaGravity (systemDynamicsVSMPC.cpp:265-272), POSITION_TRAJECTORY -> the four CoM / RPY
    // trajectories (costsVSMPC.cpp:47-68)
    std::string fileA, fileP;
    if (!glue::trajectoryFileOfGroup(ptr, "TRAJECTORY_MANAGER", fileA) || !glue::trajectoryFileOfGroup(ptr, "POSITION_TRAJECTORY", fileP))
        return false;
    std::map<std::string, std::vector<double>> va, vp;
    std::map<std::string, int> da, dp;
    if (!glue::loadTrajectoryFile(fileA, va, da, p.alphaFps) || !glue::loadTrajectoryFile(fileP, vp, dp, p.trajFps))
    {
        yError() << "Unable to load the trajectory from file.";
        return false;
    }
    if (!va.count("alphaGravity") || da["alphaGravity"] != 1)
    {
        yError() << "vsmpc::paramsFromHandler: 'alphaGravity' (1 x n) not found in " << fileA;
        return false;
    }
    p.alphaGravity = va["alphaGravity"];
    const char* keys[4] = {"positionCoM", "velocityCoM", "RPY", "RPYDot"};
    std::vector<double>* dst[4] = {&p.positionCoM, &p.velocityCoM, &p.RPY, &p.RPYDot};
    for (int k = 0; k < 4; ++k)
    {
        if (!vp.count(keys[k]) || dp[keys[k]] != 3)
        {
            yError() << "vsmpc::paramsFromHandler: '" << keys[k] << "' (3 x n) not found in " << fileP;
            return false;
        }
        *dst[k] = vp[keys[k]];
    }
    return true;
}

// what update(QPInput&) reads, field by field (the VSMPC_PK_* rows of vsmpc.h cite every getter)
inline bool fillPack(QPInput& qpInput, const std::vector<int>& sel, Pack& pk)
{
    std::shared_ptr<Robot> robot = qpInput.getRobot();
    std::shared_ptr<Robot> robotRef = qpInput.getRobotReference();
    if (robot == nullptr || robotRef == nullptr || sel.size() != VSMPC_NJ)
        return false;
    const int nJets = static_cast<int>(robot->getNJets());
    if (nJets != VSMPC_NT)
    {
        yError() << "vsmpc::fillPack: the path is built for " << VSMPC_NT << " jets (VSconstant.h:7)";
        return false;
    }
    const iDynTree::Transform wHb = robot->getBasePose();
    const iDynTree::Vector3 rpy = wHb.getRotation().asRPY();
    const iDynTree::Twist baseVel = robot->getBaseVel();
    for (int i = 0; i < 3; ++i)
    {
        for (int j = 0; j < 3; ++j)
            pk[VSMPC_PK_WRB + 3 * i + j] = wHb.getRotation()(i, j);
        pk[VSMPC_PK_OMEGA_WORLD + i] = baseVel.getAngularVec3()(i);
        pk[VSMPC_PK_RPY + i] = rpy(i);
        pk[VSMPC_PK_GRAVITY + i] = robotRef->getGravity()(i);
        pk[VSMPC_PK_BASE_POS + i] = wHb.getPosition()(i);
        pk[VSMPC_PK_P_COM + i] = robot->getPositionCoM()(i);
    }
    pk[VSMPC_PK_MASS] = robotRef->getTotalMass(); // a float member in the reference (Robot.h:338): already rounded
    for (int i = 0; i < 6; ++i)
    {
        for (int j = 0; j < 6; ++j)
            pk[VSMPC_PK_MB + 6 * i + j] = robot->getMassMatrix()(i, j);
        pk[VSMPC_PK_MOMENTUM_BODY + i] = robot->getMomentum(true)(i);
        for (int j = 0; j < VSMPC_NT; ++j)
            pk[VSMPC_PK_AMOM_BODY + VSMPC_NT * i + j] = robotRef->getMatrixAmomJets(true)(i, j);
    }
    const std::vector<iDynTree::Direction> axes = robotRef->getMatrixOfJetAxes();
    const std::vector<Eigen::Vector3d>& arms = robotRef->getMatrixOfJetArms();
    const std::vector<Eigen::MatrixXd>& Jrel = robotRef->getRelativeJacobianJetsBodyFrame();
    const std::vector<std::string>& jets = robotRef->getJetsList();
    for (int k = 0; k < VSMPC_NT; ++k)
    {
        const Eigen::MatrixXd Jk = robotRef->getJacobian(jets[k]); // 6 x (6 + nJoints)
        for (int i = 0; i < 3; ++i)
        {
            pk[VSMPC_PK_JET_AXES + 3 * k + i] = axes[k](i);
            pk[VSMPC_PK_JET_ARMS + 3 * k + i] = arms[k](i);
            for (int a = 0; a < VSMPC_NJ; ++a)
            {
                pk[VSMPC_PK_J_REL_ANG + (3 * k + i) * VSMPC_NJ + a] = Jrel[k](3 + i, sel[a]);      // bottomRows(3)
                pk[VSMPC_PK_J_JET_LIN + (3 * k + i) * VSMPC_NJ + a] = Jk(i, 6 + sel[a]);            // topRightCorner(3, nJ)
            }
        }
        pk[VSMPC_PK_THRUST + k] = robot->getJetThrusts()(k);
        pk[VSMPC_PK_THRUST_DOT_EST + k] = qpInput.getEstimatedThrustDot()(k);
        pk[VSMPC_PK_THRUST_DES + k] = qpInput.getThrustDesMPC()(k);
        pk[VSMPC_PK_THRUST_DOT_DES + k] = qpInput.getThrustDotDesMPC()(k);
        pk[VSMPC_PK_THROTTLE_PREV + k] = qpInput.getThrottleMPC()(k);
    }
    for (int i = 0; i < 3; ++i)
        for (int a = 0; a < VSMPC_NJ; ++a)
            pk[VSMPC_PK_J_COM + i * VSMPC_NJ + a] = robotRef->getJacobianCoM()(i, 6 + sel[a]);
    for (int a = 0; a < VSMPC_NJ; ++a)
        pk[VSMPC_PK_Q_CMD + a] = qpInput.getOutputQPJointsPosition()(sel[a]);
    return true;
}

// ---- the reference's class surface over libvsmpc ------------------------------------------------------------------------
class VariableSamplingMPCOnGpu
{
public:
    // IMPCProblem::configure (IMPCProblem.h:35, IMPCProblem.cpp:3-148)
    const bool configure(ParametersHandlerWeakPtr parametersHandler, QPInput& qpInput, int device = 0)
    {
        Params p;
        std::vector<std::string> names;
        std::shared_ptr<JetModel> jm = qpInput.getJetModel();
        if (!paramsFromHandler(parametersHandler, jm.get(), p, names))
            return false;
        std::shared_ptr<Robot> robot = qpInput.getRobot();
        if (robot == nullptr)
            return false;
        if (static_cast<int>(names.size()) != VSMPC_NJ)
        { // variableSamplingMPC.cpp:18-23
            yError() << "VariableSamplingMPC::setCostAndConstraints: The number of controlled joints defined in the "
                        "systemDynamic.h file is different from the size of the 'controlledJoints' parameter";
            return false;
        }
        // variableSamplingMPC.cpp:25-37: name -> index in the robot's joint list
        m_sel.clear();
        for (const std::string& n : names)
            for (size_t j = 0; j < robot->getNJoints(); ++j)
                if (n == robot->getJointName(static_cast<int>(j)))
                    m_sel.push_back(static_cast<int>(j));
        if (m_sel.size() != VSMPC_NJ)
        {
            yError() << "vsmpc: controlled joints not found in the robot's joint list";
            return false;
        }
        std::vector<double> jointPos(robot->getNJoints());
        for (size_t j = 0; j < jointPos.size(); ++j)
            jointPos[j] = robot->getJointPos()(static_cast<Eigen::Index>(j));
        Pack pk;
        if (!fillPack(qpInput, m_sel, pk) || !m_impl.configure(p, pk, jointPos, m_sel, device))
            return false;
        publish(qpInput);
        return true;
    }
    // IMPCProblem::update (IMPCProblem.h:44, IMPCProblem.cpp:150-194)
    const bool update(QPInput& qpInput)
    {
        Pack pk;
        if (!fillPack(qpInput, m_sel, pk) || !m_impl.update(pk))
            return false;
        publish(qpInput);
        return true;
    }
    const bool solveMPC() { return m_impl.solveMPC(); } // variableSamplingMPC.cpp:88-112

    // getters, variableSamplingMPC.cpp:114-227 (size mismatch -> false)
    const bool getJointsReferencePosition(Eigen::Ref<Eigen::VectorXd> v) { return m_impl.getJointsReferencePosition(v); }
    const bool getThrottleReference(Eigen::Ref<Eigen::VectorXd> v) { return m_impl.getThrottleReference(v); }
    const bool getThrustReference(Eigen::Ref<Eigen::VectorXd> v) { return m_impl.getThrustReference(v); }
    const bool getThrustDotReference(Eigen::Ref<Eigen::VectorXd> v) { return m_impl.getThrustDotReference(v); }
    const bool getFinalCoMPosition(Eigen::Ref<Eigen::VectorXd> v) { return m_impl.getFinalCoMPosition(v); }
    const bool getFinalLinMom(Eigen::Ref<Eigen::VectorXd> v) { return m_impl.getFinalLinMom(v); }
    const bool getFinalRPY(Eigen::Ref<Eigen::VectorXd> v) { return m_impl.getFinalRPY(v); }
    const bool getFinalAngMom(Eigen::Ref<Eigen::VectorXd> v) { return m_impl.getFinalAngMom(v); }
    const bool getMPCSolution(Eigen::Ref<Eigen::VectorXd> v) { return m_impl.getMPCSolution(v); }
    double getNStatesMPC() const { return m_impl.getNStatesMPC(); }
    double getNInputMPC() const { return m_impl.getNInputMPC(); }
    // inherited from IMPCProblem (IMPCProblem.h:56-118); matrices as row-major vectors of n*n / m*n
    const bool getSolution(Eigen::Ref<Eigen::VectorXd> v) { return m_impl.getSolution(v); }
    const bool getGradient(Eigen::Ref<Eigen::VectorXd> v) { return m_impl.getGradient(v); }
    const bool getLowerBound(Eigen::Ref<Eigen::VectorXd> v) { return m_impl.getLowerBound(v); }
    const bool getUpperBound(Eigen::Ref<Eigen::VectorXd> v) { return m_impl.getUpperBound(v); }
    const bool getHessian(Eigen::Ref<Eigen::VectorXd> v) { return m_impl.getHessian(v); }
    const bool getLinearConstraintMatrix(Eigen::Ref<Eigen::VectorXd> v) { return m_impl.getLinearConstraintMatrix(v); }
    const unsigned int getNOptimizationVariables() const { return static_cast<unsigned int>(m_impl.getNOptimizationVariables()); }
    const unsigned int getNConstraints() const { return static_cast<unsigned int>(m_impl.getNConstraints()); }
    // 0 solved (the reference: OsqpEigen::Status::Solved), 1 iteration cap, 2 numerical
    int getQPProblemStatus() const { return m_impl.getQPProblemStatus(); }
    VariableSamplingMPC& impl() { return m_impl; }

private:
    void publish(QPInput& qpInput)
    {
        Eigen::Vector3d pos, rpy;
        Eigen::Vector6d mom;
        m_impl.getPosCoMReference(pos);
        m_impl.getRPYReference(rpy);
        m_impl.getMomentumReference(mom);
        qpInput.setPosCoMReference(pos);             // costsVSMPC.cpp:155
        qpInput.setRPYReference(rpy);                // :156
        qpInput.setMomentumReference(mom);           // :157-160
        qpInput.setAlphaGravity(m_impl.getAlphaGravity()); // systemDynamicsVSMPC.cpp:310
    }
    VariableSamplingMPC m_impl; // vsmpc::VariableSamplingMPC of vsmpc_adapter.hpp (one instance over the C-ABI)
    std::vector<int> m_sel;
};

} // namespace vsmpc
#endif // VSMPC_REFERENCE_GLUE_HPP
