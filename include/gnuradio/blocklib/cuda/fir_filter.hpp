// gr::cuda::fir_filter<IN_T> -- decimating FIR with real taps (ccf: IN_T = gr_complex, fff:
// IN_T = float).  y[m] = sum_k h[k] x[m*D - k], phase 0.
//
// Rate-changing, so it derives gr::block (block.hpp:24-104) and sets n_consumed = n_out*D,
// n_produced = n_out itself (block_work_io.hpp:21,36); the reference scheduler has no
// history()/relative_rate, so the (ntaps-1)-sample history lives in the C-ABI
// handle on the device and every presented item can be consumed.
// An adjacent downstream multiply_const can be fused with set_fused_multiply_const(k).
#pragma once
#include <gnuradio/blocklib/cuda/cuda_block.hpp>

#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace gr {
namespace cuda {

template <class IN_T>
class fir_filter : public block, public stream_owner
{
public:
    typedef std::shared_ptr<fir_filter> sptr;
    static sptr make(size_t decimation, const std::vector<float>& taps)
    {
        auto ptr = std::make_shared<fir_filter>(decimation, taps);
        ptr->add_port(port<IN_T>::make("input", port_direction_t::INPUT));
        ptr->add_port(port<IN_T>::make("output", port_direction_t::OUTPUT));
        return ptr;
    }
    fir_filter(size_t decimation, const std::vector<float>& taps)
        : block("fir_filter"), d_decim(decimation), d_taps(taps)
    {
        build();
    }
    ~fir_filter() override { b200_fir_destroy(d_fir); }

    // fuse a downstream multiply_const (complex k for ccf, real part used for fff)
    void set_fused_multiply_const(gr_complex k)
    {
        d_fuse = true;
        d_k = k;
        build();
    }
    void set_taps(const std::vector<float>& taps)
    {
        d_taps = taps;
        build();
    }
    // Time-segment sharding (BASELINE config 5): the (ntaps-1) items that precede this block's
    // segment, read from DEVICE memory -- typically the tail of the left neighbour GPU's resident segment, a
    // peer copy over NVLink on this block's stream (b200_enable_peer_access makes it a direct one).
    void set_history_device(const void* d_hist)
    {
        check(b200_stream_activate(d_stream), "cuda::fir_filter");
        check(b200_fir_set_history(d_fir, d_hist, d_stream), "cuda::fir_filter");
    }
    int algorithm() const { return b200_fir_algorithm(d_fir); }
    std::vector<float> taps() const { return d_taps; }
    bool has_fused_multiply_const() const { return d_fuse; }
    size_t decimation() const { return d_decim; }

    work_return_code_t work(std::vector<block_work_input>& work_input,
                            std::vector<block_work_output>& work_output) override
    {
        // produce at most what fits downstream
        int64_t n_in = std::min<int64_t>(work_input[0].n_items, (int64_t)work_output[0].n_items * (int64_t)d_decim);
        int64_t nc = 0, np = 0;
        if (n_in >= (int64_t)d_decim) {
            static const bool trace = getenv("B200_TRACE_FIR") != nullptr;
            auto t0 = std::chrono::steady_clock::now();
            work_guard g(work_input, work_output, d_stream);
            auto t1 = std::chrono::steady_clock::now();
            check(b200_fir_run(d_fir, work_input[0].buffer->read_ptr(), work_output[0].buffer->write_ptr(),
                               n_in, &nc, &np, d_stream),
                  "cuda::fir_filter");
            if (trace) {
                auto t2 = std::chrono::steady_clock::now();
                std::fprintf(stderr, "fir work: t=%.1f us n_in=%lld guard %.1f us run %.1f us in=%p out=%p\n",
                             std::chrono::duration<double, std::micro>(t0.time_since_epoch()).count(), (long long)n_in,
                             std::chrono::duration<double, std::micro>(t1 - t0).count(),
                             std::chrono::duration<double, std::micro>(t2 - t1).count(), work_input[0].buffer->read_ptr(),
                             work_output[0].buffer->write_ptr());
            }
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
    void build()
    {
        if (d_fir)
            b200_fir_destroy(d_fir);
        d_fir = nullptr;
        b200_fir_params p{};
        p.taps = d_taps.data();
        p.n_taps = (int32_t)d_taps.size();
        p.decimation = (int32_t)d_decim;
        p.is_complex = std::is_same<IN_T, gr_complex>::value ? 1 : 0;
        p.fuse_multiply_const = d_fuse ? 1 : 0;
        p.k_re = d_k.real();
        p.k_im = d_k.imag();
        p.algorithm = 0;
        check(b200_fir_create(&p, &d_fir), "cuda::fir_filter");
    }
    size_t d_decim;
    std::vector<float> d_taps;
    bool d_fuse = false;
    gr_complex d_k{ 1.f, 0.f };
    b200_fir* d_fir = nullptr;
};

typedef fir_filter<gr_complex> fir_filter_ccf;
typedef fir_filter<float> fir_filter_fff;

} // namespace cuda
} // namespace gr
