// gr::cuda::fft -- N-point complex FFT per item with optional window / shift (fft_vcc semantics,
// SURVEY.md 8c).  make(fft_size, forward, window, shift) like GNU Radio's fft_vcc.
//
//   * input port:  vlen = N items, or -- with stream_input = true -- a plain complex stream from
//     which N items are consumed per transform (the FIR -> multiply_const -> FFT chain of
//     BASELINE config 3 connects an 8-byte stream to a 32 KiB-vector block; SURVEY.md 7.3
//     "Stream->vector item-size change");
//   * output port: complex vlen N, or float vlen N when a downstream complex_to_mag is fused
//     (set at make time: output = MAG / MAG_SQUARED);
//   * an upstream multiply_const_cc can be fused with pre_multiply_const.
#pragma once
#include <gnuradio/blocklib/cuda/cuda_block.hpp>

namespace gr {
namespace cuda {

enum class fft_output_t { COMPLEX = B200_FFT_OUT_COMPLEX, MAG = B200_FFT_OUT_MAG, MAG_SQUARED = B200_FFT_OUT_MAG_SQUARED };

class fft : public block, public stream_owner
{
public:
    typedef std::shared_ptr<fft> sptr;
    static sptr make(size_t fft_size, bool forward, const std::vector<float>& window = {}, bool shift = false,
                     fft_output_t output = fft_output_t::COMPLEX, bool stream_input = false,
                     bool fuse_pre_multiply_const = false, gr_complex k = gr_complex(1.f, 0.f))
    {
        auto ptr = std::make_shared<fft>(fft_size, forward, window, shift, output, stream_input,
                                         fuse_pre_multiply_const, k);
        if (stream_input)
            ptr->add_port(port<gr_complex>::make("input", port_direction_t::INPUT));
        else
            ptr->add_port(port<gr_complex>::make("input", port_direction_t::INPUT, { fft_size }));
        if (output == fft_output_t::COMPLEX)
            ptr->add_port(port<gr_complex>::make("output", port_direction_t::OUTPUT, { fft_size }));
        else
            ptr->add_port(port<float>::make("output", port_direction_t::OUTPUT, { fft_size }));
        return ptr;
    }
    fft(size_t fft_size, bool forward, const std::vector<float>& window, bool shift, fft_output_t output,
        bool stream_input, bool fuse_pre, gr_complex k)
        : block("fft"), d_n(fft_size), d_in_per_vec(stream_input ? fft_size : 1), d_forward(forward),
          d_window(window), d_shift(shift), d_output(output), d_stream_input(stream_input), d_fuse_pre(fuse_pre),
          d_k(k)
    {
        if (!window.empty() && window.size() != fft_size)
            throw std::invalid_argument("fft: window must have fft_size entries");
        b200_fft_params p{};
        p.n = (int32_t)fft_size;
        p.forward = forward ? 1 : 0;
        p.window = window.empty() ? nullptr : window.data();
        p.shift = shift ? 1 : 0;
        p.output = (int32_t)output;
        p.fuse_pre_multiply_const = fuse_pre ? 1 : 0;
        p.k_re = k.real();
        p.k_im = k.imag();
        check(b200_fft_create(&p, &d_fft), "cuda::fft");
    }
    ~fft() override { b200_fft_destroy(d_fft); }

    work_return_code_t work(std::vector<block_work_input>& work_input,
                            std::vector<block_work_output>& work_output) override
    {
        int64_t nv = std::min<int64_t>((int64_t)work_input[0].n_items / (int64_t)d_in_per_vec,
                                       (int64_t)work_output[0].n_items);
        if (nv > 0) {
            work_guard g(work_input, work_output, d_stream);
            check(b200_fft_run(d_fft, work_input[0].buffer->read_ptr(), work_output[0].buffer->write_ptr(), nv,
                               d_stream),
                  "cuda::fft");
        }
        work_input[0].n_consumed = (int)(nv * (int64_t)d_in_per_vec);
        work_output[0].n_produced = (int)nv;
        return work_return_code_t::WORK_OK;
    }
    bool done() override
    {
        synchronize();
        return block::done();
    }

    // parameters (used by the fusion pass to rebuild the block with a fused neighbour)
    size_t fft_size() const { return d_n; }
    bool forward() const { return d_forward; }
    const std::vector<float>& window() const { return d_window; }
    bool shift() const { return d_shift; }
    fft_output_t output() const { return d_output; }
    bool stream_input() const { return d_stream_input; }
    bool fused_pre_multiply_const() const { return d_fuse_pre; }
    gr_complex pre_multiply_const() const { return d_k; }

private:
    size_t d_n, d_in_per_vec;
    bool d_forward;
    std::vector<float> d_window;
    bool d_shift;
    fft_output_t d_output;
    bool d_stream_input, d_fuse_pre;
    gr_complex d_k;
    b200_fft* d_fft = nullptr;
};

} // namespace cuda
} // namespace gr
