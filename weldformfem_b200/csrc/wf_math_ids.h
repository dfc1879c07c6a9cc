// wf_math_ids.h — element-type ids shared by host and device code.
#pragma once
enum { ET_HEX8 = 0, ET_TET4 = 1, ET_QUAD4 = 2, ET_TRI3 = 3 };
