// gnuradio/types.hpp -- sample typedefs (reference runtime/include/gnuradio/types.hpp:7-16).
#pragma once
#include <complex>
#include <cstddef>
#include <cstdint>
#include <vector>

typedef std::complex<float> gr_complex;
typedef std::complex<double> gr_complexd;
typedef std::vector<int> gr_vector_int;
typedef std::vector<float> gr_vector_float;
