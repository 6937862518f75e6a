/* Stand-in for arc_utilities/pretty_print.hpp: simple_robot_models.hpp includes it and uses it only inside comments.
 * TEST INFRASTRUCTURE ONLY (see arc_helpers.hpp next to this file). */
#ifndef PRETTY_PRINT_STANDIN_HPP
#define PRETTY_PRINT_STANDIN_HPP
#include <iostream>
#include <string>
#endif
