// Shim with the reference's header name (mpc_to_line/src/MPC.h): put include/b200mpc/compat first on the include
// path and the reference's `#include "MPC.h"` resolves to the B200 library's class MPC.
#pragma once
#include "../MPC.h"
