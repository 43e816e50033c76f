// gr::cuda::multiply<T> / gr::cuda::add<T> -- two-input stream blocks on device edges
// (natural neighbours of multiply_const in real flowgraphs; same port /
// make() pattern as blocklib/blocks/include/gnuradio/blocklib/blocks/multiply_const.hpp:13-26).
#pragma once
#include <gnuradio/blocklib/cuda/cuda_block.hpp>

namespace gr {
namespace cuda {

template <class T, bool IS_ADD>
class binary_op : public sync_block, public stream_owner
{
public:
    typedef std::shared_ptr<binary_op> sptr;
    static sptr make(const size_t vlen = 1)
    {
        auto ptr = std::make_shared<binary_op>(vlen);
        ptr->add_port(port<T>::make("in0", port_direction_t::INPUT, std::vector<size_t>{ vlen }));
        ptr->add_port(port<T>::make("in1", port_direction_t::INPUT, std::vector<size_t>{ vlen }));
        ptr->add_port(port<T>::make("output", port_direction_t::OUTPUT, std::vector<size_t>{ vlen }));
        return ptr;
    }
    explicit binary_op(size_t vlen) : sync_block(IS_ADD ? "add" : "multiply"), d_vlen(vlen) {}

    work_return_code_t work(std::vector<block_work_input>& work_input,
                            std::vector<block_work_output>& work_output) override
    {
        const int n = work_output[0].n_items;
        const size_t noi = (size_t)n * d_vlen;
        {
            work_guard g(work_input, work_output, d_stream);
            check(launch(work_output[0].buffer->write_ptr(), work_input[0].buffer->read_ptr(),
                         work_input[1].buffer->read_ptr(), noi),
                  IS_ADD ? "cuda::add" : "cuda::multiply");
        }
        work_output[0].n_produced = n;
        return work_return_code_t::WORK_OK;
    }
    bool done() override
    {
        synchronize();
        return sync_block::done();
    }

private:
    int launch(void* out, const void* a, const void* b, size_t noi)
    {
        if (std::is_same<T, gr_complex>::value)
            return IS_ADD ? b200_add_cc(out, a, b, noi, d_stream) : b200_multiply_cc(out, a, b, noi, d_stream);
        return IS_ADD ? b200_add_ff((float*)out, (const float*)a, (const float*)b, noi, d_stream)
                      : b200_multiply_ff((float*)out, (const float*)a, (const float*)b, noi, d_stream);
    }
    size_t d_vlen;
};

typedef binary_op<float, false> multiply_ff;
typedef binary_op<gr_complex, false> multiply_cc;
typedef binary_op<float, true> add_ff;
typedef binary_op<gr_complex, true> add_cc;

} // namespace cuda
} // namespace gr
