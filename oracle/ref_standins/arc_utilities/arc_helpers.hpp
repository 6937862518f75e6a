/*
 * Stand-in for arc_utilities/arc_helpers.hpp (UM-ARM-Lab/arc_utilities is not vendored in the reference and not present in
 * this container) with the ONE name the reference's simple_uncertainty_models.hpp uses from it:
 * arc_helpers::TruncatedNormalDistribution (simple_uncertainty_models.hpp:25, :53), plus the two small helpers
 * simple_robot_models.hpp takes from it (UNUSED, RetrieveOrDefault).  TEST INFRASTRUCTURE ONLY - it exists so
 * that the reference header can be compiled UNMODIFIED into oracle/_ref/libref_unc.so and the oracle's sensor / actuator
 * arithmetic (everything the reference does with a draw) can be pinned against it.
 *
 * The sampling itself is what cannot be reproduced here (upstream's sampler is un-pinned and consumes an unknown number of
 * generator outputs per draw), so this stand-in does not sample: its "generator" is a tape of standard truncated normals
 * (z in [-2, 2] standard deviations) and operator() returns mean + z * stddev clamped to [lower, upper] - the value a
 * truncated-normal distribution with these parameters assigns to that standard draw.
 */
#ifndef ARC_HELPERS_STANDIN_HPP
#define ARC_HELPERS_STANDIN_HPP

/* the standard headers the real arc_helpers.hpp pulls in and the reference header relies on transitively */
#include <assert.h>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <algorithm>
#include <map>
#include <memory>
#include <limits>
#include <stdexcept>
#include <string>
#include <iostream>
#include <vector>

/* arc_helpers.hpp's unused-argument marker (simple_robot_models.hpp:147, 574) */
#ifndef UNUSED
#define UNUSED(x) (void)(x)
#endif

namespace arc_helpers
{
/* the RNG type handed to GetSensorValue / GetControlValue by the pinning shim: replays standard draws in order */
struct StandardDrawTape
{
    const double* draws;
    size_t size, next;
    double Next()
    {
        assert(next < size);
        return draws[next++];
    }
};

/* map lookup with a default (simple_robot_models.hpp:1496) */
template <typename Key, typename Value>
inline Value RetrieveOrDefault(const std::map<Key, Value>& map, const Key& key, const Value& default_value)
{
    const typename std::map<Key, Value>::const_iterator found = map.find(key);
    return (found != map.end()) ? found->second : default_value;
}

class TruncatedNormalDistribution
{
protected:
    double mean_, stddev_, lower_, upper_;

public:
    TruncatedNormalDistribution(const double mean, const double stddev, const double lower_bound, const double upper_bound)
        : mean_(mean), stddev_(stddev), lower_(lower_bound), upper_(upper_bound) {}

    template <typename Generator>
    double operator()(Generator& generator)
    {
        const double value = mean_ + generator.Next() * stddev_;
        return std::max(lower_, std::min(upper_, value));
    }
};
}

#endif
