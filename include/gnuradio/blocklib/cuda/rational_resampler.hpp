// gr::cuda::rational_resampler<IN_T> / gr::cuda::interp_fir_filter<IN_T> -- polyphase
// interpolating FIR and rational resampler with real taps (ccf: IN_T = gr_complex, fff: float).
//   y[m] = sum_k h[k] xu[m*D - k],  xu = x zero-stuffed by L      
//
// Rate-changing, so it derives gr::block (block.hpp:24-104) and reports n_consumed = groups*D,
// n_produced = groups*L itself (block_work_io.hpp:21,36).  The reference scheduler has no
// relative_rate()/history(): the ceil(T/L)-1 samples of history live in the C-ABI
// handle on the device and every call starts at phase 0.
#pragma once
#include <gnuradio/blocklib/cuda/cuda_block.hpp>

namespace gr {
namespace cuda {

template <class IN_T>
class rational_resampler : public block, public stream_owner
{
public:
    typedef std::shared_ptr<rational_resampler> sptr;
    static sptr make(size_t interpolation, size_t decimation, const std::vector<float>& taps)
    {
        auto ptr = std::make_shared<rational_resampler>(interpolation, decimation, taps);
        ptr->add_port(port<IN_T>::make("input", port_direction_t::INPUT));
        ptr->add_port(port<IN_T>::make("output", port_direction_t::OUTPUT));
        return ptr;
    }
    rational_resampler(size_t interpolation, size_t decimation, const std::vector<float>& taps)
        : block("rational_resampler"), d_interp(interpolation), d_decim(decimation), d_taps(taps)
    {
        b200_resampler_params p{};
        p.taps = d_taps.data();
        p.n_taps = (int32_t)d_taps.size();
        p.interpolation = (int32_t)d_interp;
        p.decimation = (int32_t)d_decim;
        p.is_complex = std::is_same<IN_T, gr_complex>::value ? 1 : 0;
        check(b200_resampler_create(&p, &d_rs), "cuda::rational_resampler");
    }
    ~rational_resampler() override { b200_resampler_destroy(d_rs); }

    size_t interpolation() const { return d_interp; }
    size_t decimation() const { return d_decim; }
    std::vector<float> taps() const { return d_taps; }

    work_return_code_t work(std::vector<block_work_input>& work_input,
                            std::vector<block_work_output>& work_output) override
    {
        // whole groups of D inputs -> L outputs, at most what fits downstream
        int64_t groups = std::min<int64_t>(work_input[0].n_items / (int64_t)d_decim,
                                           work_output[0].n_items / (int64_t)d_interp);
        int64_t nc = 0, np = 0;
        if (groups > 0) {
            work_guard g(work_input, work_output, d_stream);
            check(b200_resampler_run(d_rs, work_input[0].buffer->read_ptr(), work_output[0].buffer->write_ptr(),
                                     groups * (int64_t)d_decim, &nc, &np, d_stream),
                  "cuda::rational_resampler");
        }
        work_input[0].n_consumed = (int)nc;
        work_output[0].n_produced = (int)np;
        return work_return_code_t::WORK_OK;
    }
    bool done() override
    {
        synchronize();
        return block::done();
    }

private:
    size_t d_interp, d_decim;
    std::vector<float> d_taps;
    b200_resampler* d_rs = nullptr;
};

// interp_fir_filter = rational_resampler with decimation 1 (make(interpolation, taps))
template <class IN_T>
class interp_fir_filter : public rational_resampler<IN_T>
{
public:
    typedef std::shared_ptr<interp_fir_filter> sptr;
    static sptr make(size_t interpolation, const std::vector<float>& taps)
    {
        auto ptr = std::make_shared<interp_fir_filter>(interpolation, taps);
        ptr->add_port(port<IN_T>::make("input", port_direction_t::INPUT));
        ptr->add_port(port<IN_T>::make("output", port_direction_t::OUTPUT));
        return ptr;
    }
    interp_fir_filter(size_t interpolation, const std::vector<float>& taps)
        : rational_resampler<IN_T>(interpolation, 1, taps)
    {
    }
};

typedef rational_resampler<gr_complex> rational_resampler_ccf;
typedef rational_resampler<float> rational_resampler_fff;
typedef interp_fir_filter<gr_complex> interp_fir_filter_ccf;
typedef interp_fir_filter<float> interp_fir_filter_fff;

} // namespace cuda
} // namespace gr
