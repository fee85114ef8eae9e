/* Hand-written stand-in for the header GTSAM's CMake generates from
 * cmake/dllexport.h.in (reference: cmake/dllexport.h.in:31-60), Linux branch.
 * Test infrastructure only (see oracle/Makefile). */
#pragma once
#define GTSAM_EXPORT
#define GTSAM_EXTERN_EXPORT extern
