// bm_flowgraph -- flowgraph-level benchmark of the BASELINE configs on device-resident edges,
// timed the way the reference times its own benchmarks: wall clock around fg->start()/fg->wait()
// printed as [PROFILE_TIME]seconds[PROFILE_TIME] (schedulers/mt/bench/bm_copy.cpp:145-154,
// schedulers/mt/bench/cuda/bm_copy.cpp:106-114).  Flags follow bm_copy.cpp:35-57: --samples --veclen
// --nblocks --nthreads --buffer_size --rt_prio --cpus (comma separated), plus --config / --ntaps /
// --fused / --gpus.
//
//   --config 1  cuda::null_source -> fir_filter_ccf(ntaps) -> null_sink
//   --config 2  cuda::null_source -> fft(veclen, Blackman-Harris)[+ fused |.|] -> null_sink
//   --config 3  cuda::null_source -> fir_filter_ccf(1024 taps, decim 4)[+ fused k] -> fft -> null_sink
//   --config 4  cuda::null_source -> pfb_channelizer_ccf(64 channels x 16 taps) -> null_sink   (--fused 2: tensor-core DFT)
//   --config 0  cuda::null_source -> nblocks x cuda::copy -> null_sink       (bm_mt_cuda_copy shape)
//   --config 5  BASELINE configs[4]: one stream cut into --gpus time segments, each resident on its GPU
//               (cuda::vector_source), fir_filter_ccf(ntaps, default 4096) per GPU with the (ntaps-1)-sample
//               halo peer-copied from the left neighbour's segment -> null_sink
//   --config 10 BASELINE configs[0] as written: vector_source (host std::vector) -> fir_filter_ccf(ntaps)
//               -> vector_sink (host), H2D / D2H staging edges, single mt scheduler
// --gpus N (configs 0-3): N replicas of the chain, one per GPU, in ONE process under ONE mt scheduler
// (every block, ring, stream and event on its own device; SURVEY.md 8e "single process, 8 devices").
// --samples is per GPU; the JSON line reports the aggregate rate.
#include <gnuradio/blocklib/blocks/null_sink.hpp>
#include <gnuradio/blocklib/blocks/vector_sink.hpp>
#include <gnuradio/blocklib/blocks/vector_source.hpp>
#include <gnuradio/blocklib/cuda/complex_to_mag.hpp>
#include <gnuradio/blocklib/cuda/copy.hpp>
#include <gnuradio/blocklib/cuda/fft.hpp>
#include <gnuradio/blocklib/cuda/fir_filter.hpp>
#include <gnuradio/blocklib/cuda/multiply_const.hpp>
#include <gnuradio/blocklib/cuda/null_source.hpp>
#include <gnuradio/blocklib/cuda/pfb_channelizer.hpp>
#include <gnuradio/blocklib/cuda/vector_source.hpp>
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

struct options {
    int config = 2, veclen = 4096, nblocks = 4, ntaps = 0, fused = 1, clear = 0, gpus = 1, nthreads = 0, warm = 1;
    bool rt_prio = false;
    uint64_t samples = 1ull << 27;
    size_t buffer_size = 256u << 20;
    std::vector<unsigned int> cpus;
};

int main(int argc, char** argv)
{
    options opt;
    for (int i = 1; i < argc; i++) {
        std::string k = argv[i];
        if (k == "--rt_prio") {
            opt.rt_prio = true;
            continue;
        }
        if (i + 1 >= argc)
            break;
        const char* v = argv[++i];
        if (k == "--config") opt.config = atoi(v);
        else if (k == "--samples") opt.samples = strtoull(v, nullptr, 10);
        else if (k == "--veclen") opt.veclen = atoi(v);
        else if (k == "--nblocks") opt.nblocks = atoi(v);
        else if (k == "--ntaps") opt.ntaps = atoi(v);
        else if (k == "--fused") opt.fused = atoi(v);
        else if (k == "--clear") opt.clear = atoi(v);
        else if (k == "--buffer_size") opt.buffer_size = strtoull(v, nullptr, 10);
        else if (k == "--gpus") opt.gpus = atoi(v);
        else if (k == "--nthreads") opt.nthreads = atoi(v);
        else if (k == "--warm") opt.warm = atoi(v);
        else if (k == "--cpus") {
            for (const char* p = v; *p;) {
                opt.cpus.push_back((unsigned)strtoul(p, (char**)&p, 10));
                if (*p == ',')
                    p++;
            }
        }
    }
    if (opt.ntaps == 0)
        opt.ntaps = opt.config == 5 ? 4096 : 64;
    int ndev = 0;
    b200_device_count(&ndev);
    if (opt.gpus < 1 || opt.gpus > ndev) {
        std::fprintf(stderr, "bm_flowgraph: --gpus %d but %d device(s) visible\n", opt.gpus, ndev);
        return 3;
    }

    // One run.  The timed run is preceded by a short untimed one of the same flowgraph (--warm 0 disables
    // it): CUDA loads a kernel's module at its first launch, and for a run of a few milliseconds those
    // one-off loads (they grow with the size of the library, not with the work) would be most of the wall clock.
    auto run = [&](uint64_t samples, bool quiet) -> int {
        const int config = opt.config, veclen = opt.veclen;
        auto fg = flowgraph::make();
        auto sched = schedulers::scheduler_mt::make("sched", 32768);
        std::vector<std::vector<block_sptr>> chains; // per GPU, in stream order
        std::vector<std::shared_ptr<blocks::null_sink>> sinks;
        std::shared_ptr<blocks::vector_sink_c> vsnk;
        uint64_t expect_items = 0;

        if (config == 10) {
            std::vector<gr_complex> data(samples);
            for (size_t i = 0; i < data.size(); i++)
                data[i] = gr_complex((float)((i * 2654435761u) & 0xffff) / 32768.f - 1.f,
                                     (float)((i * 40503u) & 0xffff) / 32768.f - 1.f);
            auto src = blocks::vector_source_c::make(data);
            std::vector<float> taps(opt.ntaps, 1.0f / opt.ntaps);
            auto f = cuda::fir_filter_ccf::make(1, taps);
            vsnk = blocks::vector_sink_c::make(1, samples);
            fg->connect(src, 0, f, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(H2D, opt.buffer_size));
            fg->connect(f, 0, vsnk, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2H, opt.buffer_size));
            chains.push_back({ src, f, vsnk });
            expect_items = samples;
        } else {
            std::vector<std::shared_ptr<cuda::vector_source_c>> resident(opt.gpus);
            std::vector<std::shared_ptr<cuda::fir_filter_ccf>> seg_fir(opt.gpus);
            for (int g = 0; g < opt.gpus; g++) {
                if (b200_set_device(g) != B200_OK) {
                    std::fprintf(stderr, "bm_flowgraph: %s\n", b200_last_error());
                    return 3;
                }
                auto dev = [&](edge_sptr e) { e->set_custom_buffer(DEVICE_BUFFER_ARGS_ON(D2D, opt.buffer_size, g)); };
                std::vector<block_sptr> chain;
                std::shared_ptr<blocks::null_sink> snk;
                uint64_t items = 0;
                if (config == 2) {
                    auto src = cuda::null_source::make(veclen * sizeof(gr_complex), samples / veclen, opt.clear != 0);
                    auto w = blackman_harris(veclen);
                    if (opt.fused) {
                        auto f = cuda::fft::make(veclen, true, w, false, cuda::fft_output_t::MAG);
                        snk = blocks::null_sink::make(veclen * sizeof(float));
                        dev(fg->connect(src, 0, f, 0));
                        dev(fg->connect(f, 0, snk, 0));
                        chain = { src, f, snk };
                    } else {
                        auto f = cuda::fft::make(veclen, true, w);
                        auto m = cuda::complex_to_mag::make(veclen);
                        snk = blocks::null_sink::make(veclen * sizeof(float));
                        dev(fg->connect(src, 0, f, 0));
                        dev(fg->connect(f, 0, m, 0));
                        dev(fg->connect(m, 0, snk, 0));
                        chain = { src, f, m, snk };
                    }
                    items = samples / veclen;
                } else if (config == 1) {
                    auto src = cuda::null_source::make(sizeof(gr_complex), samples, opt.clear != 0);
                    std::vector<float> taps(opt.ntaps, 1.0f / opt.ntaps);
                    auto f = cuda::fir_filter_ccf::make(1, taps);
                    snk = blocks::null_sink::make(sizeof(gr_complex));
                    dev(fg->connect(src, 0, f, 0));
                    dev(fg->connect(f, 0, snk, 0));
                    chain = { src, f, snk };
                    items = samples;
                } else if (config == 5) {
                    resident[g] = cuda::vector_source_c::make_zeros(samples);
                    std::vector<float> taps(opt.ntaps, 1.0f / opt.ntaps);
                    seg_fir[g] = cuda::fir_filter_ccf::make(1, taps);
                    snk = blocks::null_sink::make(sizeof(gr_complex));
                    dev(fg->connect(resident[g], 0, seg_fir[g], 0));
                    dev(fg->connect(seg_fir[g], 0, snk, 0));
                    chain = { resident[g], seg_fir[g], snk };
                    items = samples;
                } else if (config == 4) {
                    // BASELINE configs[3]: 64-channel polyphase channelizer (16 taps per channel); --fused 2 selects the
                    // form with the DFT across branches as a tensor-core GEMM (b200_pfb_set_algorithm 2)
                    auto src = cuda::null_source::make(sizeof(gr_complex), samples, opt.clear != 0);
                    std::vector<float> taps(64 * 16);
                    for (int i = 0; i < 64 * 16; i++) { // windowed-sinc prototype
                        const double t = (i - (64 * 16 - 1) / 2.0) / 64;
                        const double sc = std::fabs(t) < 1e-12 ? 1.0 : std::sin(M_PI * t) / (M_PI * t);
                        taps[i] = (float)(sc * (0.54 - 0.46 * std::cos(2 * M_PI * i / (64 * 16 - 1))) / 64);
                    }
                    auto ch = cuda::pfb_channelizer_ccf::make(64, taps, 0, 0, opt.fused == 2 ? 2 : 1);
                    snk = blocks::null_sink::make(64 * sizeof(gr_complex));
                    dev(fg->connect(src, 0, ch, 0));
                    dev(fg->connect(ch, 0, snk, 0));
                    chain = { src, ch, snk };
                    items = samples / 64;
                } else if (config == 3) {
                    auto src = cuda::null_source::make(sizeof(gr_complex), samples, opt.clear != 0);
                    std::vector<float> taps(1024, 1.0f / 1024);
                    auto f = cuda::fir_filter_ccf::make(4, taps);
                    auto w = blackman_harris(veclen);
                    snk = blocks::null_sink::make(veclen * sizeof(gr_complex));
                    dev(fg->connect(src, 0, f, 0));
                    if (opt.fused) {
                        f->set_fused_multiply_const(gr_complex(0.5f, -0.25f));
                        auto t = cuda::fft::make(veclen, true, w, false, cuda::fft_output_t::COMPLEX, true);
                        dev(fg->connect(f, 0, t, 0));
                        dev(fg->connect(t, 0, snk, 0));
                        chain = { src, f, t, snk };
                    } else {
                        auto mul = cuda::multiply_const_cc::make(gr_complex(0.5f, -0.25f));
                        auto t = cuda::fft::make(veclen, true, w, false, cuda::fft_output_t::COMPLEX, true);
                        dev(fg->connect(f, 0, mul, 0));
                        dev(fg->connect(mul, 0, t, 0));
                        dev(fg->connect(t, 0, snk, 0));
                        chain = { src, f, mul, t, snk };
                    }
                    items = samples / 4 / veclen;
                } else {
                    auto src = cuda::null_source::make(veclen * sizeof(gr_complex), samples / veclen, opt.clear != 0);
                    chain.push_back(src);
                    node_sptr last = src;
                    for (int b = 0; b < opt.nblocks; b++) {
                        auto c = cuda::copy::make(veclen);
                        dev(fg->connect(last, 0, c, 0));
                        chain.push_back(c);
                        last = c;
                    }
                    snk = blocks::null_sink::make(veclen * sizeof(gr_complex));
                    dev(fg->connect(last, 0, snk, 0));
                    chain.push_back(snk);
                    items = samples / veclen;
                }
                chains.push_back(chain);
                sinks.push_back(snk);
                expect_items = items;
            }
            if (config == 5) // halo: the last ntaps-1 samples of the left neighbour's resident segment, GPU to GPU
                for (int g = 1; g < opt.gpus; g++) {
                    b200_set_device(g);
                    if (b200_enable_peer_access(g - 1) != B200_OK)
                        std::fprintf(stderr, "bm_flowgraph: no peer access %d -> %d (%s): the halo copy is staged\n", g, g - 1,
                                     b200_last_error());
                    seg_fir[g]->set_history_device(resident[g - 1]->device_data() + samples - (opt.ntaps - 1));
                }
            b200_set_device(0);
        }
        if (opt.nthreads > 0) { // bm_copy.cpp:112-141: split every chain into nthreads block groups
            unsigned next_cpu = 0;
            for (auto& chain : chains) {
                const int nb = (int)chain.size(), per = std::max(1, nb / opt.nthreads);
                for (int i = 0, b = 0; i < opt.nthreads && b < nb; i++) {
                    std::vector<block_sptr> grp;
                    const int take = i == opt.nthreads - 1 ? nb - b : per;
                    for (int j = 0; j < take && b < nb; j++)
                        grp.push_back(chain[b++]);
                    std::vector<unsigned int> aff;
                    if (!opt.cpus.empty())
                        aff.push_back(opt.cpus[next_cpu++ % opt.cpus.size()]);
                    sched->add_block_group(grp, "group" + std::to_string(i), aff);
                }
            }
        } else if (!opt.cpus.empty()) {
            sched->set_thread_affinity(opt.cpus);
        }
        sched->set_rt_prio(opt.rt_prio);
        fg->set_scheduler(sched);
        fg->validate();
        for (int g = 0; g < opt.gpus; g++) {
            b200_set_device(g);
            b200_device_synchronize();
        }
        b200_set_device(0);
        int64_t l0 = b200_launch_count();
        auto t1 = std::chrono::steady_clock::now();
        fg->start();
        fg->wait();
        for (int g = 0; g < opt.gpus; g++) {
            b200_set_device(g);
            b200_device_synchronize();
        }
        b200_set_device(0);
        auto t2 = std::chrono::steady_clock::now();
        double sec = std::chrono::duration<double>(t2 - t1).count();
        uint64_t got = 0;
        bool ok = true;
        if (vsnk) {
            got = vsnk->size();
            ok = got == expect_items;
        } else
            for (auto& s : sinks) {
                got += s->n_items();
                ok &= s->n_items() == expect_items;
            }
        const uint64_t total = samples * (uint64_t)(config == 10 ? 1 : opt.gpus);
        if (!quiet) {
            std::printf("[PROFILE_TIME]%f[PROFILE_TIME]\n", sec);
            std::printf("{\"config\": %d, \"gpus\": %d, \"samples\": %llu, \"samples_per_gpu\": %llu, \"seconds\": %.6f, "
                        "\"Msamples_s\": %.1f, \"sink_items\": %llu, \"expected_items\": %llu, \"kernel_launches\": %lld, "
                        "\"fused\": %d, \"buffer_size\": %zu, \"source_clears\": %d, \"nthreads\": %d, \"ntaps\": %d}\n",
                        config, config == 10 ? 1 : opt.gpus, (unsigned long long)total, (unsigned long long)samples, sec,
                        total / sec / 1e6, (unsigned long long)got,
                        (unsigned long long)(expect_items * (uint64_t)(config == 10 ? 1 : opt.gpus)),
                        (long long)(b200_launch_count() - l0), config == 10 ? 0 : opt.fused, opt.buffer_size, opt.clear,
                        opt.nthreads, opt.ntaps);
        }
        return ok ? 0 : 2;
    };
    if (opt.warm) {
        const uint64_t unit = (uint64_t)opt.veclen * 4 * 64;
        uint64_t ws = std::min<uint64_t>(opt.samples, (uint64_t)1 << 24) / unit * unit;
        if (ws >= unit)
            (void)run(ws, true);
    }
    return run(opt.samples, false);
}
