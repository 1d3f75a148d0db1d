// TEST INFRASTRUCTURE — extern "C" driver of the reference's OWN VariableSamplingMPC, compiled by oracle/build_ref.py
// from the sources where they lie under /root/reference:
//   momentum-based-linear-mpc-lib/src/variableSamplingMPC/{variableSamplingMPC,systemDynamicsVSMPC,constraintsVSMPC,costsVSMPC}.cpp
//   momentum-based-linear-mpc-lib/src/IMPCProblem/{IMPCProblem,IQPUtilsMPC,systemDynamic}.cpp
//   utils/src/{QPInput,IQPCost,IQPConstraint,FlightControlUtils,TrajectoryManager,JetModel}.cpp
// against the stand-in headers of oracle/ref_stubs/ (Eigen, iDynTree value types, YARP logging, BLF parameter handler,
// OsqpEigen, matio are not installed in the image).  This file computes nothing of the MPC: it
//   * fills a `Robot` with the outputs of its getters (the checker's data — the iDynTree kinematics behind Robot::setState
//     is outside the MPC path) and defines those getters, since utils/src/Robot.cpp cannot be built;
//   * forwards QPInput setters, the parameter table and the trajectory arrays;
//   * calls configure / update / solveMPC exactly as src/variable_sampling_mpc.py:70,111-112 does and reads results back.
#include <Eigen/Dense>
#include <standin_blf.h>
#include <standin_idyntree.h>
#include <standin_yarp.h>
#include <matio.h>

#include <cstring>
#include <map>
#include <sstream>

#define private public // Robot's data members are written directly (see above); all other headers are already included
#include "Robot.h"
#undef private
#include "QPInput.h"
#include <OsqpEigen/OsqpEigen.h>
#include <variableSamplingMPC/variableSamplingMPC.h>

// ---- Robot getters: return the stored data (utils/src/Robot.cpp:336-588 are one-line getters of the same members) ------
namespace {
std::map<const Robot*, std::map<std::string, Eigen::MatrixXd>>& jacobians()
{
    static std::map<const Robot*, std::map<std::string, Eigen::MatrixXd>> j;
    return j;
}
std::vector<std::string> splitCsv(const char* s)
{
    std::vector<std::string> out;
    std::stringstream ss(s ? s : "");
    std::string item;
    while (std::getline(ss, item, ',')) if (!item.empty()) out.push_back(item);
    return out;
}
} // namespace

const iDynTree::Twist Robot::getBaseVel() const { return m_baseVel; }
const iDynTree::Transform Robot::getBasePose() const { return m_wHb; }
const size_t Robot::getNJoints() const { return m_nJoints; }
const size_t Robot::getNJets() const { return m_nJets; }
const double Robot::getTotalMass() const { return m_totalMass; }
Eigen::Ref<const Eigen::VectorXd> Robot::getJointPos() const { return m_jointPos; }
Eigen::Ref<const Eigen::VectorXd> Robot::getJointVel() const { return m_jointVel; }
std::string Robot::getJointName(int jointPos) const { return m_axesList[jointPos]; }
Eigen::Ref<const Eigen::VectorXd> Robot::getJetThrusts() const { return m_jetThrusts; }
const iDynTree::Vector3& Robot::getGravity() const { return m_gravity; }
Eigen::Ref<const Eigen::MatrixXd> Robot::getMassMatrix() const { return m_massMatrix; }
Eigen::Ref<const Eigen::Vector6d> Robot::getMomentum(bool inBodyCoord) const { return inBodyCoord ? m_momentumBody : m_momentum; }
Eigen::Ref<const Eigen::Vector3d> Robot::getPositionCoM() const { return m_wPcom; }
Eigen::Ref<const Eigen::MatrixXd> Robot::getCentroidalMomentumMatrix() const { return m_centroidalMomentumMatrix; }
Eigen::Ref<const Eigen::MatrixXd> Robot::getJacobianCoM() const { return m_Jcom; }
const Eigen::MatrixXd Robot::getJacobian(const std::string& frameName) { return jacobians()[this].at(frameName); }
Eigen::Ref<const Eigen::MatrixXd> Robot::getMatrixAmomJets(bool inBodyCoord) const { return inBodyCoord ? m_AmomJetsBody : m_AmomJets; }
const std::vector<iDynTree::Direction> Robot::getMatrixOfJetAxes() const { return m_matrixOfJetAxes; }
const std::vector<Eigen::MatrixXd>& Robot::getRelativeJacobianJetsBodyFrame() const { return m_J_jets_body_frame; }
const std::vector<Eigen::Vector3d>& Robot::getMatrixOfJetArms() const { return m_matrixOfJetArms; }
const std::vector<std::string>& Robot::getAxesList() const { return m_axesList; }
const std::vector<std::string>& Robot::getJetsList() const { return m_jetsList; }

#ifdef VSMPC_WITH_GLUE
// the product's reference-side binding (include/vsmpc_reference_glue.hpp), compiled against the same stand-in headers and the
// reference's own QPInput.cpp: oracle/build_ref.build_glue() -> oracle/_ref/libvsmpc_reference_glue.so (links libvsmpc.so)
#include <vsmpc_reference_glue.hpp>
#endif

// ---- the handle ---------------------------------------------------------------------------------------------------------
struct RefMpc
{
#ifdef VSMPC_WITH_GLUE
    vsmpc::VariableSamplingMPCOnGpu gpu;   // driven on the SAME QPInput object as the reference class below
#endif
    std::shared_ptr<Robot> robot = std::make_shared<Robot>();
    std::shared_ptr<BipedalLocomotion::ParametersHandler::YarpImplementation> params
        = std::make_shared<BipedalLocomotion::ParametersHandler::YarpImplementation>();
    std::map<std::string, std::shared_ptr<BipedalLocomotion::ParametersHandler::YarpImplementation>> groups;
    QPInput qpInput;
    VariableSamplingMPC mpc;
    int nJ = 0, nJets = 0;
};

// the trajectory files of the matio stand-in, by file name
static std::map<std::string, mat_t*>& matFiles()
{
    static std::map<std::string, mat_t*> f;
    return f;
}
mat_t* standinMatOpen(const char* path)
{
    auto it = matFiles().find(path);
    if (it == matFiles().end()) return nullptr;
    it->second->next = 0;
    return it->second;
}

extern "C" {

void ref_mpc_mat_variable(const char* file, const char* var, const double* data, int d0, int d1)
{
    mat_t*& m = matFiles()[file];
    if (!m) m = new mat_t;
    matvar_t* v = nullptr;
    for (matvar_t* w : m->vars) if (w->name_store == var) v = w;
    if (!v) { v = new matvar_t; m->vars.push_back(v); }
    v->name_store = var;
    v->data_store.assign(data, data + size_t(d0) * size_t(d1));       // column-major d0 x d1, as MATLAB stores it
    v->dims_store[0] = size_t(d0); v->dims_store[1] = size_t(d1);
    v->name = const_cast<char*>(v->name_store.c_str());
    v->data = v->data_store.data();
    v->dims = v->dims_store;
}

void* ref_mpc_create(const char* jointNamesCsv, const char* jetNamesCsv)
{
    RefMpc* h = new RefMpc;
    Robot& r = *h->robot;
    r.m_axesList = splitCsv(jointNamesCsv);
    r.m_jetsList = splitCsv(jetNamesCsv);
    h->nJ = int(r.m_axesList.size());
    h->nJets = int(r.m_jetsList.size());
    r.m_nJoints = size_t(h->nJ);
    r.m_nJets = size_t(h->nJets);
    r.m_nExtWrenches = 0;
    r.m_jointPos = Eigen::VectorXd::Zero(h->nJ);
    r.m_jointVel = Eigen::VectorXd::Zero(h->nJ);
    r.m_jetThrusts = Eigen::VectorXd::Zero(h->nJets);
    h->qpInput.setRobot(h->robot);                                      // src/variable_sampling_mpc.py:49-53
    h->qpInput.setRobotReference(h->robot);
    h->qpInput.setVectorsCollectionServer(std::make_shared<BipedalLocomotion::YarpUtilities::VectorsCollectionServer>());
    h->qpInput.setJetModel(std::make_shared<JetModel>());
    return h;
}

void ref_mpc_destroy(void* p)
{
    RefMpc* h = static_cast<RefMpc*>(p);
    jacobians().erase(h->robot.get());
    delete h;
}

static BipedalLocomotion::ParametersHandler::IParametersHandler* handlerOf(RefMpc* h, const char* group)
{
    if (!group || !*group) return h->params.get();
    auto& g = h->groups[group];
    if (!g) {
        g = std::make_shared<BipedalLocomotion::ParametersHandler::YarpImplementation>();
        h->params->setGroup(group, g);
    }
    return g.get();
}

void ref_mpc_param_numbers(void* p, const char* group, const char* name, const double* v, int n)
{
    handlerOf(static_cast<RefMpc*>(p), group)->setNumbers(name, std::vector<double>(v, v + n));
}

void ref_mpc_param_strings(void* p, const char* group, const char* name, const char* csv)
{
    handlerOf(static_cast<RefMpc*>(p), group)->setStrings(name, splitCsv(csv));
}

// All matrices row-major.  J_rel_body: nJets x 6 x nJ, J_jet_lin: nJets x 3 x nJ (the joint columns of the linear rows of
// each jet frame's Jacobian), J_com: 3 x nJ (the joint columns of the CoM Jacobian).
void ref_mpc_set_robot(void* p, const double* wRb, const double* basePos, const double* omegaWorld, const double* Mb,
                       const double* pCom, const double* momentumBody, const double* AmomBody, const double* jetAxes,
                       const double* jetArms, const double* JrelBody, const double* JjetLin, const double* Jcom,
                       const double* thrust, const double* jointPos, const double* gravity)
{
    RefMpc* h = static_cast<RefMpc*>(p);
    Robot& r = *h->robot;
    const int nJ = h->nJ, nJets = h->nJets;
    iDynTree::Rotation R;
    iDynTree::Position pos;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) R(i, j) = wRb[i * 3 + j];
        pos(i) = basePos[i];
        r.m_baseVel.getAngularVec3()(i) = omegaWorld[i];
        r.m_gravity(i) = gravity[i];
        r.m_wPcom(i) = pCom[i];
    }
    r.m_wHb.setRotation(R);
    r.m_wHb.setPosition(pos);
    // mass matrix: the 6x6 base block is data; the joint block (read only by the dead code of
    // systemDynamicsVSMPC.cpp:117-126) is the identity so that its inverses exist
    r.m_massMatrix = Eigen::MatrixXd::Identity(6 + nJ, 6 + nJ);
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) r.m_massMatrix(i, j) = Mb[i * 6 + j];
    r.m_totalMass = float(Mb[0]);                                       // Robot.cpp:332 (a float member)
    for (int i = 0; i < 6; ++i) r.m_momentumBody(i) = momentumBody[i];
    r.m_AmomJetsBody.resize(6, nJets);
    for (int i = 0; i < 6; ++i) for (int j = 0; j < nJets; ++j) r.m_AmomJetsBody(i, j) = AmomBody[i * nJets + j];
    r.m_matrixOfJetAxes.assign(size_t(nJets), iDynTree::Direction());
    r.m_matrixOfJetArms.assign(size_t(nJets), Eigen::Vector3d());
    r.m_J_jets_body_frame.assign(size_t(nJets), Eigen::MatrixXd());
    for (int k = 0; k < nJets; ++k) {
        for (int i = 0; i < 3; ++i) {
            r.m_matrixOfJetAxes[size_t(k)](i) = jetAxes[k * 3 + i];
            r.m_matrixOfJetArms[size_t(k)](i) = jetArms[k * 3 + i];
        }
        Eigen::MatrixXd& Jr = r.m_J_jets_body_frame[size_t(k)];
        Jr.resize(6, nJ);
        for (int i = 0; i < 6; ++i) for (int j = 0; j < nJ; ++j) Jr(i, j) = JrelBody[(k * 6 + i) * nJ + j];
        Eigen::MatrixXd J = Eigen::MatrixXd::Zero(6, 6 + nJ);
        for (int i = 0; i < 3; ++i) for (int j = 0; j < nJ; ++j) J(i, 6 + j) = JjetLin[(k * 3 + i) * nJ + j];
        jacobians()[&r][r.m_jetsList[size_t(k)]] = J;
        r.m_jetThrusts(k) = thrust[k];
    }
    r.m_Jcom = Eigen::MatrixXd::Zero(3, 6 + nJ);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < nJ; ++j) r.m_Jcom(i, 6 + j) = Jcom[i * nJ + j];
    for (int j = 0; j < nJ; ++j) r.m_jointPos(j) = jointPos[j];
}

int ref_mpc_set_input(void* p, const char* name, const double* v, int n)
{
    RefMpc* h = static_cast<RefMpc*>(p);
    Eigen::VectorXd x = Eigen::Map<Eigen::VectorXd>(v, n);
    const std::string s(name);
    if (s == "ThrottleMPC") h->qpInput.setThrottleMPC(x);
    else if (s == "ThrustDesMPC") h->qpInput.setThrustDesMPC(x);
    else if (s == "ThrustDotDesMPC") h->qpInput.setThrustDotDesMPC(x);
    else if (s == "EstimatedThrustDot") h->qpInput.setEstimatedThrustDot(x);
    else if (s == "OutputQPJointsPosition") h->qpInput.setOutputQPJointsPosition(x);
    else return 1;
    return 0;
}

void ref_mpc_set_qp_solver(OsqpEigen::QPSolveFn f) { OsqpEigen::qpSolveFunction() = f; }

int ref_mpc_configure(void* p)
{
    RefMpc* h = static_cast<RefMpc*>(p);
    return h->mpc.configure(h->params, h->qpInput) ? 0 : 1;
}
int ref_mpc_update(void* p)
{
    RefMpc* h = static_cast<RefMpc*>(p);
    return h->mpc.update(h->qpInput) ? 0 : 1;
}
int ref_mpc_solve(void* p)
{
    RefMpc* h = static_cast<RefMpc*>(p);
    return h->mpc.solveMPC() ? 0 : 1;
}
int ref_mpc_nvar(void* p) { return int(static_cast<RefMpc*>(p)->mpc.getNOptimizationVariables()); }
int ref_mpc_ncon(void* p) { return int(static_cast<RefMpc*>(p)->mpc.getNConstraints()); }
int ref_mpc_status(void* p) { return int(static_cast<RefMpc*>(p)->mpc.getQPProblemStatus()); }

static void copyOut(const Eigen::View& m, double* out) // row-major
{
    for (Eigen::Index i = 0; i < m.rows(); ++i) for (Eigen::Index j = 0; j < m.cols(); ++j) out[i * m.cols() + j] = m(i, j);
}

// the QP as IMPCProblem holds it after update(): P (n x n), q (n), A (m x n), l, u (m); row-major
void ref_mpc_get_qp(void* p, double* P, double* q, double* A, double* l, double* u)
{
    VariableSamplingMPC& mpc = static_cast<RefMpc*>(p)->mpc;
    copyOut(mpc.getHessian(), P);
    copyOut(mpc.getGradient(), q);
    copyOut(mpc.getLinearConstraintMatrix(), A);
    copyOut(mpc.getLowerBound(), l);
    copyOut(mpc.getUpperBound(), u);
}

// getters of variableSamplingMPC.h:18-40 + the QPInput fields update() writes
void ref_mpc_get_output(void* p, double* jointsRef, double* throttle, double* thrust, double* thrustDot, double* finalState,
                        double* solution, double* qpInputOut /* alphaGravity, posCoMRef[3], rpyRef[3], momentumRef[6] */)
{
    RefMpc* h = static_cast<RefMpc*>(p);
    VariableSamplingMPC& mpc = h->mpc;
    Eigen::VectorXd j(h->nJ), t(h->nJets), T(h->nJets), Td(h->nJets), v3(3);
    mpc.getJointsReferencePosition(j); copyOut(j, jointsRef);
    mpc.getThrottleReference(t); copyOut(t, throttle);
    mpc.getThrustReference(T); copyOut(T, thrust);
    mpc.getThrustDotReference(Td); copyOut(Td, thrustDot);
    if (finalState) {
        try {       // the final state is empty until the first solved tick (the stand-in's views are range-checked)
            mpc.getFinalCoMPosition(v3); copyOut(v3, finalState + 0);
            mpc.getFinalLinMom(v3); copyOut(v3, finalState + 3);
            mpc.getFinalRPY(v3); copyOut(v3, finalState + 6);
            mpc.getFinalAngMom(v3); copyOut(v3, finalState + 9);
        } catch (const std::out_of_range&) {
            for (int i = 0; i < 12; ++i) finalState[i] = 0.0;
        }
    }
    copyOut(mpc.getSolution(), solution);
    qpInputOut[0] = h->qpInput.getAlphaGravity();
    copyOut(h->qpInput.getPosCoMReference(), qpInputOut + 1);
    copyOut(h->qpInput.getRPYReference(), qpInputOut + 4);
    copyOut(h->qpInput.getMomentumReference(), qpInputOut + 7);
}

#ifdef VSMPC_WITH_GLUE
// ---- the same call sequence on vsmpc::VariableSamplingMPCOnGpu: configure / update / solveMPC with the reference's own
// argument types (weak_ptr<IParametersHandler>, QPInput&), getters through Eigen::Ref -------------------------------------
int ref_glue_configure(void* p)
{
    RefMpc* h = static_cast<RefMpc*>(p);
    return h->gpu.configure(h->params, h->qpInput) ? 0 : 1;
}
int ref_glue_update(void* p)
{
    RefMpc* h = static_cast<RefMpc*>(p);
    return h->gpu.update(h->qpInput) ? 0 : 1;
}
int ref_glue_solve(void* p) { return static_cast<RefMpc*>(p)->gpu.solveMPC() ? 0 : 1; }
int ref_glue_status(void* p) { return static_cast<RefMpc*>(p)->gpu.getQPProblemStatus(); }
int ref_glue_nvar(void* p) { return int(static_cast<RefMpc*>(p)->gpu.getNOptimizationVariables()); }
int ref_glue_ncon(void* p) { return int(static_cast<RefMpc*>(p)->gpu.getNConstraints()); }
// same layout as ref_mpc_get_output; qpInputOut is read from the shared QPInput object (what the glue's update() wrote)
int ref_glue_get_output(void* p, double* jointsRef, double* throttle, double* thrust, double* thrustDot, double* finalState,
                        double* solution, double* qpInputOut)
{
    RefMpc* h = static_cast<RefMpc*>(p);
    Eigen::VectorXd j(h->nJ), t(h->nJets), T(h->nJets), Td(h->nJets), v3(3), z(Eigen::Index(h->gpu.getNOptimizationVariables()));
    bool ok = h->gpu.getJointsReferencePosition(j) && h->gpu.getThrottleReference(t) && h->gpu.getThrustReference(T)
              && h->gpu.getThrustDotReference(Td) && h->gpu.getSolution(z);
    copyOut(j, jointsRef); copyOut(t, throttle); copyOut(T, thrust); copyOut(Td, thrustDot); copyOut(z, solution);
    ok = ok && h->gpu.getFinalCoMPosition(v3); copyOut(v3, finalState + 0);
    ok = ok && h->gpu.getFinalLinMom(v3); copyOut(v3, finalState + 3);
    ok = ok && h->gpu.getFinalRPY(v3); copyOut(v3, finalState + 6);
    ok = ok && h->gpu.getFinalAngMom(v3); copyOut(v3, finalState + 9);
    // wrong sizes must be refused like the reference's getters (variableSamplingMPC.cpp:114-227)
    Eigen::VectorXd bad(h->nJets + 1);
    ok = ok && !h->gpu.getThrustReference(bad);
    qpInputOut[0] = h->qpInput.getAlphaGravity();
    copyOut(h->qpInput.getPosCoMReference(), qpInputOut + 1);
    copyOut(h->qpInput.getRPYReference(), qpInputOut + 4);
    copyOut(h->qpInput.getMomentumReference(), qpInputOut + 7);
    return ok ? 0 : 1;
}
// overwrite the four published QPInput fields (so that a test sees which class wrote them last)
void ref_glue_clear_published(void* p, double v)
{
    RefMpc* h = static_cast<RefMpc*>(p);
    Eigen::Vector3d a; a << v, v, v;
    Eigen::Vector6d m; m << v, v, v, v, v, v;
    h->qpInput.setPosCoMReference(a);
    h->qpInput.setRPYReference(a);
    h->qpInput.setMomentumReference(m);
    h->qpInput.setAlphaGravity(v);
}
// the pack vsmpc::fillPack builds from the QPInput / Robot getters (compared with the Python pack builder)
int ref_glue_fill_pack(void* p, const int* sel, double* pack359)
{
    RefMpc* h = static_cast<RefMpc*>(p);
    vsmpc::Pack pk;
    if (!vsmpc::fillPack(h->qpInput, std::vector<int>(sel, sel + 8), pk)) return 1;
    std::memcpy(pack359, pk.v, sizeof(pk.v));
    return 0;
}
#endif

} // extern "C"
