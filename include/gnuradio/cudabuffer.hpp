// gnuradio/cudabuffer.hpp -- source compatibility with the reference's CUDA edge buffers:
// flowgraphs written against runtime/include/gnuradio/cudabuffer.hpp:11-86 (cuda_buffer,
// cuda_buffer_type, cuda_buffer_properties, CUDA_BUFFER_ARGS_{H2D,D2H,D2D}) compile unchanged
// and get the B200 device-resident ring (gnuradio/devicebuffer.hpp).
#pragma once
#include <gnuradio/devicebuffer.hpp>

namespace gr {
using cuda_buffer_type = device_buffer_type;
using cuda_buffer_properties = device_buffer_properties;
using cuda_buffer = device_buffer;
} // namespace gr

#define CUDA_BUFFER_ARGS_H2D cuda_buffer::make, cuda_buffer_properties::make(cuda_buffer_type::H2D)
#define CUDA_BUFFER_ARGS_D2H cuda_buffer::make, cuda_buffer_properties::make(cuda_buffer_type::D2H)
#define CUDA_BUFFER_ARGS_D2D cuda_buffer::make, cuda_buffer_properties::make(cuda_buffer_type::D2D)
