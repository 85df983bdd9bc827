// Stand-in for the CMake-generated include/config.hpp.in of the reference:
// everything optional is off except METIS (the toolkit's static METIS 5 is
// linked, see oracle/Makefile).
#ifndef SCHWARZ_CONFIG_HPP_SHIM
#define SCHWARZ_CONFIG_HPP_SHIM
#define SCHW_HAVE_METIS 1
#define SCHW_HAVE_CHOLMOD 0
#define SCHW_HAVE_UMFPACK 0
#define SCHW_HAVE_DEALII 0
#define SCHW_HAVE_CUDA 0
#define SCHW_HAVE_HWLOC 0
#endif
