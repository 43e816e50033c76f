// gr::cuda -- common plumbing of the B200 GPU blocks: every block owns one CUDA stream (as the
// reference's cuda::copy does, blocklib/cuda/include/gnuradio/blocklib/cuda/copy.hpp:41), never
// synchronises it inside work(), and orders itself against neighbouring blocks through the
// events of the device-resident edge buffers (gnuradio/devicebuffer.hpp).  C-ABI failures are
// turned into exceptions: the reference scheduler cannot handle WORK_ERROR
// (schedulers/mt/lib/graph_executor.cpp:103-132 never leaves its retry loop on it).
#pragma once
#include <gnuradio/devicebuffer.hpp>
#include <gnuradio/sync_block.hpp>

#include <b200dsp.h>

namespace gr {
namespace cuda {

inline void check(int rc, const char* what)
{
    if (rc != B200_OK)
        throw std::runtime_error(std::string(what) + ": " + b200_last_error());
}

// A GPU block lives on the device that is current when it is constructed: its stream, its C-ABI handles
// (taps, history, twiddles) and -- through work_guard, which activates the stream's device -- every launch
// of its work().  `b200_set_device(k)` before make() places a block on GPU k; the reference has no device
// notion at all, its cuda::copy takes whatever device is current (copy.cpp:41-63).
class stream_owner
{
protected:
    b200_stream_t d_stream = nullptr;
    int d_device = 0;

public:
    stream_owner()
    {
        check(b200_get_device(&d_device), "get_device");
        check(b200_stream_create(&d_stream), "stream_create");
    }
    virtual ~stream_owner()
    {
        if (d_stream) {
            b200_stream_synchronize(d_stream);
            b200_stream_destroy(d_stream);
        }
    }
    int device() const { return d_device; }
    stream_owner(const stream_owner&) = delete;
    b200_stream_t stream() const { return d_stream; }
    void synchronize() { check(b200_stream_synchronize(d_stream), "stream_synchronize"); }
};

typedef device_stream_guard<std::vector<block_work_input>, std::vector<block_work_output>> work_guard;

} // namespace cuda
} // namespace gr
