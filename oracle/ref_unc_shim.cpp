/*
 * ref_unc_shim.cpp - extern "C" wrapper around the reference's OWN, UNMODIFIED simple_uncertainty_models.hpp
 * (TruncatedNormalUncertainSensor :20-46, TruncatedNormalUncertainVelocityActuator :48-106), compiled against the stand-in
 * arc_utilities headers of oracle/ref_standins (the only external name the header needs is
 * arc_helpers::TruncatedNormalDistribution; the stand-in maps a replayed standard draw to the distribution's value).
 * Built by oracle/Makefile into oracle/_ref/libref_unc.so from the header where it lies under $(REFERENCE)/include; used by
 * tests/test_oracle_unc.py to pin the oracle's SensorValue / Actuator against the reference.  TEST INFRASTRUCTURE ONLY.
 */
#include <uncertainty_planning_core/simple_uncertainty_models.hpp>

extern "C" {

/* out[i] = TruncatedNormalUncertainSensor(lower, upper).GetSensorValue(process_values[i], rng) with rng replaying standard_draws */
void ref_sensor_values(double noise_lower_bound, double noise_upper_bound, int n, const double* process_values, const double* standard_draws, double* out)
{
    const simple_uncertainty_models::TruncatedNormalUncertainSensor sensor(noise_lower_bound, noise_upper_bound);
    arc_helpers::StandardDrawTape tape = {standard_draws, (size_t)n, 0};
    for (int i = 0; i < n; i++) out[i] = sensor.GetSensorValue(process_values[i], tape);
}

/* out[i] = TruncatedNormalUncertainVelocityActuator(noise_bound, limit).GetControlValue(inputs[i], rng); noiseless[i] = GetControlValue(inputs[i]) */
void ref_actuator_values(double noise_bound, double actuator_limit, int n, const double* inputs, const double* standard_draws, double* out, double* noiseless)
{
    const simple_uncertainty_models::TruncatedNormalUncertainVelocityActuator actuator(noise_bound, actuator_limit);
    arc_helpers::StandardDrawTape tape = {standard_draws, (size_t)n, 0};
    for (int i = 0; i < n; i++)
    {
        out[i] = actuator.GetControlValue(inputs[i], tape);
        noiseless[i] = actuator.GetControlValue(inputs[i]);
    }
}

double ref_actuator_max_velocity_noise(double noise_bound, double actuator_limit)
{
    const simple_uncertainty_models::TruncatedNormalUncertainVelocityActuator actuator(noise_bound, actuator_limit);
    return actuator.GetMaxVelocityNoise();
}

}
