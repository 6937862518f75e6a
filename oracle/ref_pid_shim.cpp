/*
 * ref_pid_shim.cpp - extern "C" wrapper around the reference's OWN, UNMODIFIED simple_pid_controller.hpp
 * (the only reference file that compiles without ROS / Eigen / arc_utilities / sdf_tools).
 * Built by oracle/Makefile into oracle/_ref/libref_pid.so from the header where it lies under
 * $(REFERENCE)/include; used by tests/ to pin the oracle's Pid and to generate tests/golden/pid_*.json.
 * TEST INFRASTRUCTURE ONLY.
 */
#include <uncertainty_planning_core/simple_pid_controller.hpp>

extern "C" double ref_pid_sequence(double kp, double ki, double kd, double clamp, const double* errors, const double* dts,
                                   int n, double* out_terms)
{
    simple_pid_controller::SimplePIDController pid(kp, ki, kd, clamp);
    double last = 0.0;
    for (int i = 0; i < n; i++)
    {
        last = pid.ComputeFeedbackTerm(errors[i], dts[i]);
        if (out_terms) out_terms[i] = last;
    }
    return last;
}
