// gr::cuda::complex_to_mag -- out = sqrtf(re^2 + im^2), complex64 -> float32, vlen scalars per
// item (block absent from the reference snapshot; follows the blocklib pattern
// of blocklib/blocks/include/gnuradio/blocklib/blocks/multiply_const.hpp:13-26).
#pragma once
#include <gnuradio/blocklib/cuda/cuda_block.hpp>

namespace gr {
namespace cuda {

class complex_to_mag : public sync_block, public stream_owner
{
public:
    typedef std::shared_ptr<complex_to_mag> sptr;
    static sptr make(const size_t vlen = 1, bool squared = false)
    {
        auto ptr = std::make_shared<complex_to_mag>(vlen, squared);
        ptr->add_port(port<gr_complex>::make("input", port_direction_t::INPUT, std::vector<size_t>{ vlen }));
        ptr->add_port(port<float>::make("output", port_direction_t::OUTPUT, std::vector<size_t>{ vlen }));
        return ptr;
    }
    complex_to_mag(size_t vlen, bool squared) : sync_block("complex_to_mag"), d_vlen(vlen), d_squared(squared) {}

    work_return_code_t work(std::vector<block_work_input>& work_input,
                            std::vector<block_work_output>& work_output) override
    {
        const int n = work_output[0].n_items;
        {
            work_guard g(work_input, work_output, d_stream);
            auto fn = d_squared ? b200_complex_to_mag_squared : b200_complex_to_mag;
            check(fn((float*)work_output[0].buffer->write_ptr(), work_input[0].buffer->read_ptr(),
                     (size_t)n * d_vlen, d_stream),
                  "cuda::complex_to_mag");
        }
        work_output[0].n_produced = n;
        return work_return_code_t::WORK_OK;
    }
    bool done() override
    {
        synchronize();
        return sync_block::done();
    }

    size_t vlen() const { return d_vlen; }
    bool squared() const { return d_squared; }

private:
    size_t d_vlen;
    bool d_squared;
};

} // namespace cuda
} // namespace gr
