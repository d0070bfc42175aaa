// Unity build of the reference oracle (TEST INFRASTRUCTURE).
//
// Every reference translation unit is #included from where it lies under /root/reference/src (never copied),
// after all standard and shim headers. With -DORACLE_FP64 the token `float` is redefined to `double` AFTER the
// system headers, which turns the single-precision reference into an exact FP64 restatement of the same
// arithmetic: every f-suffixed literal on the hot path is dyadic (0.5f 0.25f 0.75f 2.0f 3.0f 4.0f 8.0f 16.0f)
// and pi always appears as (float)M_PI.
// `private`/`protected` are opened so the harness can call the per-item calculators
// (e.g. SU2FrgCore::_calculateVertexTwoParticle, src/SU2/SU2FrgCore.hpp:60) on a bounded item sample.
#define _USE_MATH_DEFINES
#include <vector>
#include <functional>
#include <thread>
#include <mutex>
#include <iostream>
#include <sstream>
#include <fstream>
#include <string>
#include <exception>
#include <stdexcept>
#include <algorithm>
#include <numeric>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <math.h>
#include <array>
#include <tuple>
#include <map>
#include <set>
#include <list>
#include <istream>
#include <ostream>
#include <iomanip>
#include <chrono>
#include <regex>
#include <filesystem>
#include <optional>
#include <memory>
#include <limits>
#include <type_traits>
#include <utility>
#include <omp.h>

#ifdef ORACLE_FP64
#define H5SHIM_REAL double
#endif
#include <boost/date_time.hpp>
#include <boost/regex.hpp>
#include <boost/filesystem.hpp>
#include <boost/format.hpp>
#include <boost/property_tree/ptree.hpp>
#include <boost/property_tree/xml_parser.hpp>
#include <hdf5.h>

#ifdef ORACLE_FP64
#define float double
#define powf pow
#endif
#define private public
#define protected public

#include "FrgCommon.cpp"
#include "lib/Log.cpp"
#include "Measurement.cpp"
#include "LatticeModelFactory.cpp"
#undef PI
#undef __EPSILON
#include "TaskFileParser.cpp"
#ifdef HARNESS_B200_FACTORY
// the product's factory (spinparser_b200/host/FrgCoreFactory_b200.cpp) instead of the reference's: same harness, same
// reference host code, flow cores running on the GPU through libpffrg -- the end-to-end drop-in test
#include "FrgCoreFactory_b200.cpp"
#else
#include "FrgCoreFactory.cpp"
#endif
#include "SU2/SU2FrgCore.cpp"
#include "SU2/SU2MeasurementCorrelation.cpp"
#include "XYZ/XYZFrgCore.cpp"
#include "XYZ/XYZMeasurementCorrelation.cpp"
#include "TRI/TRIFrgCore.cpp"
#include "TRI/TRIMeasurementCorrelation.cpp"

#include "ref_harness.cpp"
