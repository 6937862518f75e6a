/* Stand-in for arc_utilities/eigen_helpers.hpp: simple_uncertainty_models.hpp includes it (line 13) but uses nothing from it.
 * TEST INFRASTRUCTURE ONLY (see arc_helpers.hpp next to this file). */
#ifndef EIGEN_HELPERS_STANDIN_HPP
#define EIGEN_HELPERS_STANDIN_HPP
#endif
