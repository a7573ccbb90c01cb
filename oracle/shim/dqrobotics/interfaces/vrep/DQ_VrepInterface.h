// TEST INFRASTRUCTURE — forwards to the dqrobotics stand-in (see ../DQ.h).
#pragma once
#include <dqrobotics/DQ.h>
