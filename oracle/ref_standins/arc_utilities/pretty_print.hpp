/* Stand-in for arc_utilities/pretty_print.hpp: the reference's headers use PrettyPrint::PrettyPrint only in debug strings
 * (commented-out prints of simple_robot_models.hpp, UncertaintyPlannerState::Print).  TEST INFRASTRUCTURE ONLY (see arc_helpers.hpp
 * next to this file). */
#ifndef PRETTY_PRINT_STANDIN_HPP
#define PRETTY_PRINT_STANDIN_HPP
#include <iostream>
#include <string>

namespace PrettyPrint
{
template <typename T>
inline std::string PrettyPrint(const T&, const bool add_delimiters = false, const std::string& separator = ", ")
{
    (void)add_delimiters;
    (void)separator;
    return std::string("<value>");
}
}
#endif
