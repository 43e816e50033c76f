// gr::cuda::multiply_const<T> -- out = in * k over n_items*vlen scalars on the device.
// Same make(k, vlen) and typedef suffixes as the CPU block
// (blocklib/blocks/include/gnuradio/blocklib/blocks/multiply_const.hpp:13-40) and the reference
// CUDA block (blocklib/cuda/include/gnuradio/blocklib/cuda/multiply_const.hpp:8-43, float only,
// whose k is never initialised -- SURVEY.md 2.2 K2).
#pragma once
#include <gnuradio/blocklib/cuda/cuda_block.hpp>

namespace gr {
namespace cuda {

template <class T>
class multiply_const : public sync_block, public stream_owner
{
public:
    typedef std::shared_ptr<multiply_const> sptr;
    static sptr make(const T k, const size_t vlen = 1)
    {
        auto ptr = std::make_shared<multiply_const>(k, vlen);
        ptr->add_port(port<T>::make("input", port_direction_t::INPUT, std::vector<size_t>{ vlen }));
        ptr->add_port(port<T>::make("output", port_direction_t::OUTPUT, std::vector<size_t>{ vlen }));
        return ptr;
    }
    multiply_const(T k, size_t vlen) : sync_block("multiply_const"), d_k(k), d_vlen(vlen) {}

    work_return_code_t work(std::vector<block_work_input>& work_input,
                            std::vector<block_work_output>& work_output) override
    {
        const int n = work_output[0].n_items;
        const size_t noi = (size_t)n * d_vlen;
        {
            work_guard g(work_input, work_output, d_stream);
            check(launch(work_output[0].buffer->write_ptr(), work_input[0].buffer->read_ptr(), noi),
                  "cuda::multiply_const");
        }
        work_output[0].n_produced = n;
        return work_return_code_t::WORK_OK;
    }
    bool done() override
    {
        synchronize();
        return sync_block::done();
    }
    T k() const { return d_k; }
    size_t vlen() const { return d_vlen; }

private:
    int launch(void* out, const void* in, size_t noi);
    T d_k;
    size_t d_vlen;
};

template <>
inline int multiply_const<float>::launch(void* out, const void* in, size_t noi)
{
    return b200_multiply_const_ff((float*)out, (const float*)in, d_k, noi, d_stream);
}
template <>
inline int multiply_const<gr_complex>::launch(void* out, const void* in, size_t noi)
{
    return b200_multiply_const_cc(out, in, d_k.real(), d_k.imag(), noi, d_stream);
}
template <>
inline int multiply_const<std::int16_t>::launch(void* out, const void* in, size_t noi)
{
    return b200_multiply_const_ss((int16_t*)out, (const int16_t*)in, d_k, noi, d_stream);
}
template <>
inline int multiply_const<std::int32_t>::launch(void* out, const void* in, size_t noi)
{
    return b200_multiply_const_ii((int32_t*)out, (const int32_t*)in, d_k, noi, d_stream);
}

typedef multiply_const<std::int16_t> multiply_const_ss;
typedef multiply_const<std::int32_t> multiply_const_ii;
typedef multiply_const<float> multiply_const_ff;
typedef multiply_const<gr_complex> multiply_const_cc;

} // namespace cuda
} // namespace gr
