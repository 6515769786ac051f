/* forwards the reference header name to the B200 host layer */
#pragma once
#include <mpl_b200/traj_solver.hpp>
