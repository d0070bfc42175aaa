// Shim (test infrastructure): lets the UNMODIFIED reference sources compile and run without libhdf5. The HDF5 calls of the reference's
// EffectiveAction checkpoint code and MeasurementCorrelation writers (e.g. SU2EffectiveAction.hpp:79-205,
// SU2MeasurementCorrelation.cpp:179-355) are served by the product's minimal HDF5 layer spinparser_b200/host/hdf5_min.hpp: an in-memory
// tree per file (the harness serialises the trees with dump_all()) that is also stored as a real HDF5 file (superblock version 0) when
// H5MIN_DISK=1 -- off by default in the oracle binaries, whose tests read the in-memory trees.
#pragma once
#ifndef H5SHIM_REAL
#define H5SHIM_REAL float
#endif
#define H5MIN_REAL H5SHIM_REAL
#ifndef H5MIN_DISK_DEFAULT
#define H5MIN_DISK_DEFAULT 0
#endif
#include "../../spinparser_b200/host/hdf5_min.hpp"
namespace h5shim = h5min;
