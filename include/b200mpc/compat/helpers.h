// Shim with the reference's header name (mpc_to_line/src/helpers.h): polyeval / polyfit from the B200 library.
#pragma once
#include "../MPC.h"
