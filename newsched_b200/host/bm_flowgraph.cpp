// bm_flowgraph -- flowgraph-level benchmark of the BASELINE configs on device-resident edges,
// timed the way the reference times its own benchmarks: wall clock around fg->start()/fg->wait()
// printed as [PROFILE_TIME]seconds[PROFILE_TIME] (schedulers/mt/bench/bm_copy.cpp:145-154,
// schedulers/mt/bench/cuda/bm_copy.cpp:106-114).  Flag names follow bm_copy.cpp:35-57 where they
// apply (--samples, --veclen, --nblocks, --buffer_size).
//
//   --config 1  cuda::null_source -> fir_filter_ccf(ntaps) -> null_sink
//   --config 2  cuda::null_source -> fft(veclen, Blackman-Harris)[+ fused |.|] -> null_sink
//   --config 3  cuda::null_source -> fir_filter_ccf(1024 taps, decim 4)[+ fused k] -> fft -> null_sink
//   --config 0  cuda::null_source -> nblocks x cuda::copy -> null_sink       (bm_mt_cuda_copy shape)
//   --config 10 BASELINE configs[0] as written: vector_source (host std::vector) -> fir_filter_ccf(ntaps)
//               -> vector_sink (host), H2D / D2H staging edges, single mt scheduler
#include <gnuradio/blocklib/blocks/null_sink.hpp>
#include <gnuradio/blocklib/blocks/vector_sink.hpp>
#include <gnuradio/blocklib/blocks/vector_source.hpp>
#include <gnuradio/blocklib/cuda/complex_to_mag.hpp>
#include <gnuradio/blocklib/cuda/copy.hpp>
#include <gnuradio/blocklib/cuda/fft.hpp>
#include <gnuradio/blocklib/cuda/fir_filter.hpp>
#include <gnuradio/blocklib/cuda/multiply_const.hpp>
#include <gnuradio/blocklib/cuda/null_source.hpp>
#include <gnuradio/devicebuffer.hpp>
#include <gnuradio/flowgraph.hpp>
#include <gnuradio/schedulers/mt/scheduler_mt.hpp>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

using namespace gr;

static std::vector<float> blackman_harris(int N)
{
    std::vector<float> w(N);
    for (int n = 0; n < N; n++) {
        double t = (double)n / (N - 1);
        w[n] = (float)(0.35875 - 0.48829 * std::cos(2 * M_PI * t) + 0.14128 * std::cos(4 * M_PI * t) -
                       0.01168 * std::cos(6 * M_PI * t));
    }
    return w;
}

int main(int argc, char** argv)
{
    int config = 2, veclen = 4096, nblocks = 4, ntaps = 64, fused = 1, clear = 0;
    uint64_t samples = 1ull << 27;
    size_t buffer_size = 256u << 20;
    for (int i = 1; i + 1 < argc; i += 2) {
        std::string k = argv[i];
        const char* v = argv[i + 1];
        if (k == "--config") config = atoi(v);
        else if (k == "--samples") samples = strtoull(v, nullptr, 10);
        else if (k == "--veclen") veclen = atoi(v);
        else if (k == "--nblocks") nblocks = atoi(v);
        else if (k == "--ntaps") ntaps = atoi(v);
        else if (k == "--fused") fused = atoi(v);
        else if (k == "--clear") clear = atoi(v);
        else if (k == "--buffer_size") buffer_size = strtoull(v, nullptr, 10);
    }
    int warm = 1;
    for (int i = 1; i + 1 < argc; i += 2)
        if (std::string(argv[i]) == "--warm")
            warm = atoi(argv[i + 1]);
    // One flowgraph run.  The timed run is preceded by a short untimed one of the same flowgraph
    // (--warm 0 disables it): CUDA loads a kernel's module at its first launch, and for a run of a
    // few milliseconds those one-off loads (they grow with the size of the library, not with the
    // work) would be most of the wall clock.
    auto run = [&](uint64_t samples, bool quiet) -> int {
    auto fg = flowgraph::make();
        auto sched = schedulers::scheduler_mt::make("sched", 32768);
        auto dev = [&](edge_sptr e) { e->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2D, buffer_size)); };
        std::shared_ptr<blocks::null_sink> snk;
        uint64_t expect_items = 0;
        if (config == 10) {
            std::vector<gr_complex> data(samples);
            for (size_t i = 0; i < data.size(); i++)
                data[i] = gr_complex((float)((i * 2654435761u) & 0xffff) / 32768.f - 1.f,
                                     (float)((i * 40503u) & 0xffff) / 32768.f - 1.f);
            auto src = blocks::vector_source_c::make(data);
            std::vector<float> taps(ntaps, 1.0f / ntaps);
            auto f = cuda::fir_filter_ccf::make(1, taps);
            auto vsnk = blocks::vector_sink_c::make(1, samples);
            fg->connect(src, 0, f, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(H2D, buffer_size));
            fg->connect(f, 0, vsnk, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2H, buffer_size));
            fg->set_scheduler(sched);
            fg->validate();
            b200_device_synchronize();
            int64_t l0 = b200_launch_count();
            auto t1 = std::chrono::steady_clock::now();
            fg->start();
            fg->wait();
            b200_device_synchronize();
            auto t2 = std::chrono::steady_clock::now();
            double sec = std::chrono::duration<double>(t2 - t1).count();
            if (!quiet) std::printf("[PROFILE_TIME]%f[PROFILE_TIME]\n", sec);
            if (!quiet) std::printf("{\"config\": 10, \"samples\": %llu, \"seconds\": %.6f, \"Msamples_s\": %.1f, \"sink_items\": %llu, "
                        "\"expected_items\": %llu, \"kernel_launches\": %lld, \"fused\": 0, \"buffer_size\": %zu, "
                        "\"source_clears\": 0}\n",
                        (unsigned long long)samples, sec, samples / sec / 1e6, (unsigned long long)vsnk->data().size(),
                        (unsigned long long)samples, (long long)(b200_launch_count() - l0), buffer_size);
            return vsnk->data().size() == samples ? 0 : 2;
        }
        if (config == 2) {
            auto src = cuda::null_source::make(veclen * sizeof(gr_complex), samples / veclen, clear != 0);
            auto w = blackman_harris(veclen);
            if (fused) {
                auto f = cuda::fft::make(veclen, true, w, false, cuda::fft_output_t::MAG);
                snk = blocks::null_sink::make(veclen * sizeof(float));
                dev(fg->connect(src, 0, f, 0));
                dev(fg->connect(f, 0, snk, 0));
            } else {
                auto f = cuda::fft::make(veclen, true, w);
                auto m = cuda::complex_to_mag::make(veclen);
                snk = blocks::null_sink::make(veclen * sizeof(float));
                dev(fg->connect(src, 0, f, 0));
                dev(fg->connect(f, 0, m, 0));
                dev(fg->connect(m, 0, snk, 0));
            }
            expect_items = samples / veclen;
        } else if (config == 1) {
            auto src = cuda::null_source::make(sizeof(gr_complex), samples, clear != 0);
            std::vector<float> taps(ntaps, 1.0f / ntaps);
            auto f = cuda::fir_filter_ccf::make(1, taps);
            snk = blocks::null_sink::make(sizeof(gr_complex));
            dev(fg->connect(src, 0, f, 0));
            dev(fg->connect(f, 0, snk, 0));
            expect_items = samples;
        } else if (config == 3) {
            auto src = cuda::null_source::make(sizeof(gr_complex), samples, clear != 0);
            std::vector<float> taps(1024, 1.0f / 1024);
            auto f = cuda::fir_filter_ccf::make(4, taps);
            auto w = blackman_harris(veclen);
            snk = blocks::null_sink::make(veclen * sizeof(gr_complex));
            dev(fg->connect(src, 0, f, 0));
            if (fused) {
                f->set_fused_multiply_const(gr_complex(0.5f, -0.25f));
                auto t = cuda::fft::make(veclen, true, w, false, cuda::fft_output_t::COMPLEX, true);
                dev(fg->connect(f, 0, t, 0));
                dev(fg->connect(t, 0, snk, 0));
            } else {
                auto mul = cuda::multiply_const_cc::make(gr_complex(0.5f, -0.25f));
                auto t = cuda::fft::make(veclen, true, w, false, cuda::fft_output_t::COMPLEX, true);
                dev(fg->connect(f, 0, mul, 0));
                dev(fg->connect(mul, 0, t, 0));
                dev(fg->connect(t, 0, snk, 0));
            }
            expect_items = samples / 4 / veclen;
        } else {
            auto src = cuda::null_source::make(veclen * sizeof(gr_complex), samples / veclen, clear != 0);
            node_sptr last = src;
            for (int b = 0; b < nblocks; b++) {
                auto c = cuda::copy::make(veclen);
                dev(fg->connect(last, 0, c, 0));
                last = c;
            }
            snk = blocks::null_sink::make(veclen * sizeof(gr_complex));
            dev(fg->connect(last, 0, snk, 0));
            expect_items = samples / veclen;
        }
        fg->set_scheduler(sched);
        fg->validate();
        b200_device_synchronize();
        int64_t l0 = b200_launch_count();
        auto t1 = std::chrono::steady_clock::now();
        fg->start();
        fg->wait();
        b200_device_synchronize();
        auto t2 = std::chrono::steady_clock::now();
        double sec = std::chrono::duration<double>(t2 - t1).count();
        if (!quiet) std::printf("[PROFILE_TIME]%f[PROFILE_TIME]\n", sec);
        if (!quiet) std::printf("{\"config\": %d, \"samples\": %llu, \"seconds\": %.6f, \"Msamples_s\": %.1f, \"sink_items\": %llu, "
                    "\"expected_items\": %llu, \"kernel_launches\": %lld, \"fused\": %d, \"buffer_size\": %zu, "
                    "\"source_clears\": %d}\n",
                    config, (unsigned long long)samples, sec, samples / sec / 1e6, (unsigned long long)snk->n_items(),
                    (unsigned long long)expect_items, (long long)(b200_launch_count() - l0), fused, buffer_size, clear);
        return snk->n_items() == expect_items ? 0 : 2;
    };
    if (warm) {
        const uint64_t unit = (uint64_t)veclen * 4 * 64;
        uint64_t ws = std::min<uint64_t>(samples, (uint64_t)1 << 24) / unit * unit;
        if (ws >= unit)
            (void)run(ws, true);
    }
    return run(samples, false);
}
