// CL/cl_ext.h -- empty stand-in (TEST INFRASTRUCTURE ONLY; see CL/cl2.hpp).
#pragma once
