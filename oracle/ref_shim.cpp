// TEST INFRASTRUCTURE — extern "C" access to the reference's own JetModel class
// (src/flight-controller/utils/include/JetModel.h:12-95, compiled from src/flight-controller/utils/src/JetModel.cpp
// where it lies under /root/reference by oracle/build_ref.py).  Nothing is computed here.
#include "JetModel.h"

extern "C" {
// which: 0 f, 1 g, 2 df_dT, 3 df_dTdot, 4 dg_dT, 5 dg_dTdot  (arguments: standardised thrust and thrust rate)
double ref_jet_poly(int which, double T, double Tdot)
{
    JetModel j;
    switch (which) {
    case 0: return j.compute_f(T, Tdot);
    case 1: return j.compute_g(T, Tdot);
    case 2: return j.compute_df_dT(T, Tdot);
    case 3: return j.compute_df_dTdot(T, Tdot);
    case 4: return j.compute_dg_dT(T, Tdot);
    default: return j.compute_dg_dTdot(T, Tdot);
    }
}
// which: 0 compute_v, 1 standardizeThrust, 2 standardizeThrustDot, 3 standardizeThrottle, 4 destandardizeThrust,
//        5 destandardizeThrustDot, 6 destandardizeThrottle, 7 getThrustStandardDeviation (argument ignored)
double ref_jet_scalar(int which, double x)
{
    JetModel j;
    switch (which) {
    case 0: return j.compute_v(x);
    case 1: return j.standardizeThrust_u2T(x);
    case 2: return j.standardizeThrustDot_u2T(x);
    case 3: return j.standardizeThrottle_u2T(x);
    case 4: return j.destandardizeThrust_u2T(x);
    case 5: return j.destandardizeThrustDot_u2T(x);
    case 6: return j.destandardizeThrottle_u2T(x);
    default: return j.getThrustStandardDeviation_u2T();
    }
}
}
