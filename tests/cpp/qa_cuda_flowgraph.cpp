// GPU QA at flowgraph level: the B200 blocks and the device-resident edge buffer driven by the
// thread-per-block scheduler through block::work(), exactly as the reference's own tests do
// (schedulers/mt/test/cuda/qa_scheduler_mt_cuda_copy.cpp:20-86, qa_scheduler_mt.cpp:79-135,
// qa_block_grouping.cpp:15-66).  Expected values come from the CPU oracle (liboracle.so).
#include <gnuradio/blocklib/blocks/null_sink.hpp>
#include <gnuradio/blocklib/blocks/vector_sink.hpp>
#include <gnuradio/blocklib/blocks/vector_source.hpp>
#include <gnuradio/blocklib/cuda/complex_to_mag.hpp>
#include <gnuradio/blocklib/cuda/copy.hpp>
#include <gnuradio/blocklib/cuda/fft.hpp>
#include <gnuradio/blocklib/cuda/fir_filter.hpp>
#include <gnuradio/blocklib/cuda/fusion.hpp>
#include <gnuradio/blocklib/cuda/multiply.hpp>
#include <gnuradio/blocklib/cuda/multiply_const.hpp>
#include <gnuradio/blocklib/cuda/null_source.hpp>
#include <gnuradio/blocklib/cuda/pfb_channelizer.hpp>
#include <gnuradio/blocklib/cuda/rational_resampler.hpp>
#include <gnuradio/blocklib/cuda/stream_to_vector.hpp>
#include <gnuradio/blocklib/cuda/vector_source.hpp>
#include <gnuradio/cudabuffer.hpp>
#include <gnuradio/cudabuffer_pinned.hpp>
#include <gnuradio/flowgraph.hpp>
#include <gnuradio/schedulers/mt/scheduler_mt.hpp>

#include <cmath>
#include <random>

#include "qa_common.hpp"

using namespace gr;

extern "C" {
int64_t orc_fir_ccf_f64(float*, const float*, int64_t, const float*, int, int, const float*);
int64_t orc_fir_fff_f64(float*, const float*, int64_t, const float*, int, int, const float*);
int64_t orc_resample_ccf_f64(float*, const float*, int64_t, const float*, int, int, int, const float*);
int64_t orc_resample_fff_f64(float*, const float*, int64_t, const float*, int, int, int, const float*);
int orc_fft_f64(float*, const float*, int64_t, int, int, const float*, int);
void orc_window_blackmanharris(float*, int);
void orc_multiply_const_cc(float*, const float*, float, float, int64_t);
void orc_complex_to_mag(float*, const float*, int64_t);
void orc_multiply_cc(float*, const float*, const float*, int64_t);
void orc_add_f(float*, const float*, const float*, int64_t);
int64_t orc_pfb_channelizer_f64(float*, const float*, int64_t, const float*, int, int, const float*);
}

static std::vector<gr_complex> noise(size_t n, unsigned seed)
{
    std::mt19937 g(seed);
    std::uniform_real_distribution<float> u(-1.f, 1.f);
    std::vector<gr_complex> v(n);
    for (auto& x : v)
        x = gr_complex(u(g), u(g));
    return v;
}
static std::vector<float> rtaps(size_t n, unsigned seed)
{
    std::mt19937 g(seed);
    std::uniform_real_distribution<float> u(-1.f, 1.f);
    std::vector<float> v(n);
    for (auto& x : v)
        x = u(g) / (float)n;
    return v;
}
template <class A, class B>
static double rel_rms(const std::vector<A>& a, const std::vector<B>& b)
{
    if (a.size() != b.size())
        return 1e9;
    double num = 0, den = 0;
    for (size_t i = 0; i < a.size(); i++) {
        num += std::norm(std::complex<double>(a[i]) - std::complex<double>(b[i]));
        den += std::norm(std::complex<double>(b[i]));
    }
    return den > 0 ? std::sqrt(num / den) : std::sqrt(num);
}
static const double TOL = 1e-5; // north_star: FIR / FFT within 1e-5 relative RMS

// ---- the reference's two CUDA tests, source-compatible ------------------------------------
QA_TEST(SchedulerMTTest, CudaCopyBasic)
{
    int veclen = 1024;
    int num_samples = veclen * 100;
    std::vector<gr_complex> input_data(num_samples);
    for (int i = 0; i < num_samples; i++)
        input_data[i] = gr_complex(i, -i);
    auto src = blocks::vector_source_c::make(input_data, false, veclen);
    auto snk1 = blocks::vector_sink_c::make(veclen);
    auto copy1 = cuda::copy::make(veclen);
    auto copy2 = cuda::copy::make(veclen);
    auto fg = flowgraph::make();
    fg->connect(src, 0, copy1, 0)->set_custom_buffer(CUDA_BUFFER_ARGS_H2D);
    fg->connect(copy1, 0, copy2, 0)->set_custom_buffer(CUDA_BUFFER_ARGS_D2D);
    fg->connect(copy2, 0, snk1, 0)->set_custom_buffer(CUDA_BUFFER_ARGS_D2H);
    auto sched = schedulers::scheduler_mt::make("sched", 32768);
    fg->set_scheduler(sched);
    sched->add_block_group({ copy1, copy2 });
    fg->validate();
    fg->start();
    fg->wait();
    EXPECT_EQ(snk1->data(), input_data);
}

QA_TEST(SchedulerMTTest, CudaCopyMultiThreaded)
{
    int veclen = 1024;
    int num_samples = veclen * 100;
    std::vector<gr_complex> input_data(num_samples);
    for (int i = 0; i < num_samples; i++)
        input_data[i] = gr_complex(i, -i);
    auto src = blocks::vector_source_c::make(input_data, false, veclen);
    auto snk1 = blocks::vector_sink_c::make(veclen);
    auto copy1 = cuda::copy::make(veclen);
    auto copy2 = cuda::copy::make(veclen);
    auto fg = flowgraph::make();
    fg->connect(src, 0, copy1, 0)->set_custom_buffer(CUDA_BUFFER_ARGS_H2D);
    fg->connect(copy1, 0, copy2, 0)->set_custom_buffer(CUDA_BUFFER_ARGS_D2D);
    fg->connect(copy2, 0, snk1, 0)->set_custom_buffer(CUDA_BUFFER_ARGS_D2H);
    auto sched = schedulers::scheduler_mt::make("sched", 32768);
    fg->set_scheduler(sched);
    fg->validate();
    fg->start();
    fg->wait();
    EXPECT_EQ(snk1->data(), input_data);
}

// pinned ("zero-copy") edges, as bm_mt_cuda_copy -m 1 uses them (bench/cuda/bm_copy.cpp:82-99)
QA_TEST(SchedulerMTTest, CudaCopyPinnedBuffers)
{
    auto in = noise(1500007, 41);
    auto src = blocks::vector_source_c::make(in);
    auto snk = blocks::vector_sink_c::make();
    auto c1 = cuda::copy::make(1), c2 = cuda::copy::make(1);
    auto k = cuda::multiply_const_cc::make(gr_complex(0.f, 1.f));
    auto fg = flowgraph::make();
    fg->connect(src, 0, c1, 0)->set_custom_buffer(PINNED_BUFFER_ARGS_SIZED(1u << 20));
    fg->connect(c1, 0, k, 0)->set_custom_buffer(PINNED_BUFFER_ARGS_SIZED(3u << 20));
    fg->connect(k, 0, c2, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2D, 2u << 20));
    fg->connect(c2, 0, snk, 0)->set_custom_buffer(CUDA_BUFFER_PINNED_ARGS);
    fg->set_scheduler(schedulers::scheduler_mt::make("sched", 1 << 20));
    fg->validate();
    fg->run();
    std::vector<gr_complex> exp(in.size());
    orc_multiply_const_cc((float*)exp.data(), (const float*)in.data(), 0.f, 1.f, (int64_t)in.size());
    EXPECT_EQ(snk->data().size(), exp.size());
    EXPECT_EQ(snk->data(), exp);
}

// small rings force many wrap-arounds of the doubly mapped device window
QA_TEST(DeviceBuffer, SmallRingManyWraps)
{
    auto in = noise(3000017, 1);
    auto src = blocks::vector_source_c::make(in);
    auto snk = blocks::vector_sink_c::make();
    auto c1 = cuda::copy::make(1), c2 = cuda::copy::make(1), c3 = cuda::copy::make(1);
    auto fg = flowgraph::make();
    fg->connect(src, 0, c1, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(H2D, 2u << 20));
    fg->connect(c1, 0, c2, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2D, 2u << 20));
    fg->connect(c2, 0, c3, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2D, 6u << 20));
    fg->connect(c3, 0, snk, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2H, 2u << 20));
    fg->set_scheduler(schedulers::scheduler_mt::make());
    fg->validate();
    fg->run();
    EXPECT_EQ(snk->data().size(), in.size());
    EXPECT_EQ(snk->data(), in);
}

// qa_scheduler_mt.cpp:79-135 on the GPU: fan-out of a host-fed edge into N device chains, k = 1
QA_TEST(SchedulerMTTest, BlockFanoutCuda)
{
    int num_samples = 1000000;
    std::vector<gr_complex> input_data(num_samples);
    for (int i = 0; i < num_samples; i++)
        input_data[i] = gr_complex(2 * i, 2 * i + 1);
    for (auto nblocks : { 2, 8 }) {
        auto src = blocks::vector_source_c::make(input_data);
        std::vector<std::shared_ptr<blocks::vector_sink_c>> sinks(nblocks);
        auto fg = flowgraph::make();
        for (int i = 0; i < nblocks; i++) {
            auto mult = cuda::multiply_const_cc::make(gr_complex(1.0f, 0.0f));
            sinks[i] = blocks::vector_sink_c::make();
            fg->connect(src, 0, mult, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(H2D, 4u << 20));
            fg->connect(mult, 0, sinks[i], 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2H, 4u << 20));
        }
        fg->set_scheduler(schedulers::scheduler_mt::make("sched", 8192));
        fg->validate();
        fg->run();
        for (int n = 0; n < nblocks; n++)
            EXPECT_EQ(sinks[n]->data(), input_data);
    }
}

// device-side fan-out (copy_items between device rings) + chain of 16 blocks in 4 block groups
QA_TEST(SchedulerBlockGrouping, CudaChainAndDeviceFanout)
{
    auto in = noise(1 << 20, 2);
    auto src = blocks::vector_source_c::make(in);
    auto fg = flowgraph::make();
    auto sched = schedulers::scheduler_mt::make();
    node_sptr last = src;
    bool first = true;
    for (int g = 0; g < 4; g++) {
        std::vector<block_sptr> grp;
        for (int b = 0; b < 4; b++) {
            auto m = cuda::multiply_const_cc::make(gr_complex(1.0f, 0.0f));
            auto e = fg->connect(last, 0, m, 0);
            if (first)
                e->set_custom_buffer(DEVICE_BUFFER_ARGS_H2D);
            else
                e->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2D, 8u << 20));
            first = false;
            last = m;
            grp.push_back(m);
        }
        sched->add_block_group(grp);
    }
    auto tail1 = cuda::copy::make(1);
    auto tail2 = cuda::multiply_const_cc::make(gr_complex(0.f, 1.f));
    auto snk1 = blocks::vector_sink_c::make(), snk2 = blocks::vector_sink_c::make();
    fg->connect(last, 0, tail1, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_D2D);
    fg->connect(last, 0, tail2, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_D2D);
    fg->connect(tail1, 0, snk1, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_D2H);
    fg->connect(tail2, 0, snk2, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_D2H);
    fg->set_scheduler(sched);
    fg->validate();
    fg->run();
    EXPECT_EQ(snk1->data(), in);
    std::vector<gr_complex> exp(in.size());
    orc_multiply_const_cc((float*)exp.data(), (const float*)in.data(), 0.f, 1.f, (int64_t)in.size());
    EXPECT_EQ(snk2->data(), exp);
}

// BASELINE config 1: vector_source -> fir_filter_ccf (64 taps) -> vector_sink
QA_TEST(Config1, FirCcf64)
{
    auto in = noise(1 << 21, 3);
    auto taps = rtaps(64, 4);
    for (size_t ring : { (size_t)1 << 20, (size_t)64 << 20 }) {
        auto src = blocks::vector_source_c::make(in);
        auto fir = cuda::fir_filter_ccf::make(1, taps);
        auto snk = blocks::vector_sink_c::make();
        auto fg = flowgraph::make();
        fg->connect(src, 0, fir, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(H2D, ring));
        fg->connect(fir, 0, snk, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2H, ring));
        fg->set_scheduler(schedulers::scheduler_mt::make());
        fg->validate();
        fg->run();
        std::vector<gr_complex> exp(in.size());
        orc_fir_ccf_f64((float*)exp.data(), (const float*)in.data(), (int64_t)in.size(), taps.data(), 64, 1, nullptr);
        EXPECT_EQ(snk->data().size(), exp.size());
        EXPECT_TRUE(rel_rms(snk->data(), exp) < TOL);
    }
}

QA_TEST(Config1, FirFffDecim)
{
    auto inc = noise(700001, 5);
    std::vector<float> in(inc.size());
    for (size_t i = 0; i < in.size(); i++)
        in[i] = inc[i].real();
    auto taps = rtaps(129, 6);
    auto src = blocks::vector_source_f::make(in);
    auto fir = cuda::fir_filter_fff::make(5, taps);
    auto snk = blocks::vector_sink_f::make();
    auto fg = flowgraph::make();
    fg->connect(src, 0, fir, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(H2D, 1u << 20));
    fg->connect(fir, 0, snk, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2H, 1u << 20));
    fg->set_scheduler(schedulers::scheduler_mt::make());
    fg->validate();
    fg->run();
    std::vector<float> exp(in.size() / 5);
    orc_fir_fff_f64(exp.data(), in.data(), (int64_t)in.size(), taps.data(), 129, 5, nullptr);
    EXPECT_EQ(snk->data().size(), exp.size());
    EXPECT_TRUE(rel_rms(snk->data(), exp) < TOL);
}

// interpolating FIR and rational resampler as rate-changing blocks on device edges
// full-rate float stream, 64 taps: the tensor-core form of fir_filter_fff (two 4096-sample runs per tile) inside a
// flowgraph; small rings, so work() windows start on arbitrary 4-byte boundaries and cross many tiles
QA_TEST(Config1, FirFff64TensorCore)
{
    auto inc = noise(900007, 15);
    std::vector<float> in(inc.size());
    for (size_t i = 0; i < in.size(); i++)
        in[i] = inc[i].imag();
    auto taps = rtaps(64, 16);
    auto src = blocks::vector_source_f::make(in);
    auto fir = cuda::fir_filter_fff::make(1, taps);
    auto snk = blocks::vector_sink_f::make();
    auto fg = flowgraph::make();
    fg->connect(src, 0, fir, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(H2D, 1u << 20));
    fg->connect(fir, 0, snk, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2H, 1u << 20));
    fg->set_scheduler(schedulers::scheduler_mt::make());
    fg->validate();
    fg->run();
    std::vector<float> exp(in.size());
    orc_fir_fff_f64(exp.data(), in.data(), (int64_t)in.size(), taps.data(), 64, 1, nullptr);
    EXPECT_EQ(snk->data().size(), exp.size());
    EXPECT_TRUE(rel_rms(snk->data(), exp) < TOL);
}

QA_TEST(Resampler, InterpAndRational)
{
    auto in = noise(300000, 11);
    {
        auto taps = rtaps(96, 12);
        auto src = blocks::vector_source_c::make(in);
        auto rs = cuda::interp_fir_filter_ccf::make(4, taps);
        auto snk = blocks::vector_sink_c::make();
        auto fg = flowgraph::make();
        fg->connect(src, 0, rs, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(H2D, 1u << 20));
        fg->connect(rs, 0, snk, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2H, 1u << 20));
        fg->set_scheduler(schedulers::scheduler_mt::make());
        fg->validate();
        fg->run();
        std::vector<gr_complex> exp(in.size() * 4);
        orc_resample_ccf_f64((float*)exp.data(), (const float*)in.data(), (int64_t)in.size(), taps.data(), 96, 4, 1, nullptr);
        EXPECT_EQ(snk->data().size(), exp.size());
        EXPECT_TRUE(rel_rms(snk->data(), exp) < TOL);
    }
    {
        std::vector<float> inf(in.size());
        for (size_t i = 0; i < inf.size(); i++)
            inf[i] = in[i].imag();
        auto taps = rtaps(211, 13);
        auto src = blocks::vector_source_f::make(inf);
        auto rs = cuda::rational_resampler_fff::make(3, 7, taps);
        auto snk = blocks::vector_sink_f::make();
        auto fg = flowgraph::make();
        fg->connect(src, 0, rs, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(H2D, 1u << 20));
        fg->connect(rs, 0, snk, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2H, 1u << 20));
        fg->set_scheduler(schedulers::scheduler_mt::make());
        fg->validate();
        fg->run();
        std::vector<float> exp(inf.size() / 7 * 3);
        orc_resample_fff_f64(exp.data(), inf.data(), (int64_t)inf.size(), taps.data(), 211, 3, 7, nullptr);
        EXPECT_EQ(snk->data().size(), exp.size());
        EXPECT_TRUE(rel_rms(snk->data(), exp) < TOL);
    }
}

// BASELINE config 2 with real data: source -> fft(4096, Blackman-Harris) -> complex_to_mag -> sink,
// once as two blocks over a device edge and once with the |.| fused into the FFT epilogue
QA_TEST(Config2, FftMag)
{
    const int N = 4096, nv = 300;
    auto in = noise((size_t)N * nv, 7);
    std::vector<float> w(N);
    orc_window_blackmanharris(w.data(), N);
    std::vector<gr_complex> X(in.size());
    orc_fft_f64((float*)X.data(), (const float*)in.data(), nv, N, 1, w.data(), 0);
    std::vector<float> exp(in.size());
    for (size_t i = 0; i < in.size(); i++)
        exp[i] = (float)std::abs(std::complex<double>(X[i]));
    for (int fused = 0; fused < 2; fused++) {
        auto src = blocks::vector_source_c::make(in, false, N);
        auto snk = blocks::vector_sink_f::make(N);
        auto fg = flowgraph::make();
        if (!fused) {
            auto f = cuda::fft::make(N, true, w);
            auto m = cuda::complex_to_mag::make(N);
            fg->connect(src, 0, f, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(H2D, 4u << 20));
            fg->connect(f, 0, m, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2D, 4u << 20));
            fg->connect(m, 0, snk, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2H, 4u << 20));
        } else {
            auto f = cuda::fft::make(N, true, w, false, cuda::fft_output_t::MAG);
            fg->connect(src, 0, f, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_H2D);
            fg->connect(f, 0, snk, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_D2H);
        }
        fg->set_scheduler(schedulers::scheduler_mt::make());
        fg->validate();
        fg->run();
        EXPECT_EQ(snk->data().size(), exp.size());
        EXPECT_TRUE(rel_rms(snk->data(), exp) < TOL);
    }
}

// stream -> vector -> fft (vector input) -> vector -> stream: the item-size adapters around a
// vector block; equals the stream-input fft bit for bit
QA_TEST(Adapters, StreamToVectorAndBack)
{
    const int N = 1024, nv = 200;
    auto in = noise((size_t)N * nv + 77, 17); // 77 trailing samples never fill a vector
    std::vector<float> w(N);
    orc_window_blackmanharris(w.data(), N);
    std::vector<gr_complex> exp((size_t)N * nv);
    orc_fft_f64((float*)exp.data(), (const float*)in.data(), nv, N, 1, w.data(), 0);
    auto src = blocks::vector_source_c::make(in);
    auto s2v = cuda::stream_to_vector::make(sizeof(gr_complex), N);
    auto f = cuda::fft::make(N, true, w);
    auto v2s = cuda::vector_to_stream::make(sizeof(gr_complex), N);
    auto snk = blocks::vector_sink_c::make();
    auto fg = flowgraph::make();
    fg->connect(src, 0, s2v, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(H2D, 1u << 20));
    fg->connect(s2v, 0, f, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2D, 1u << 20));
    fg->connect(f, 0, v2s, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2D, 1u << 20));
    fg->connect(v2s, 0, snk, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2H, 1u << 20));
    fg->set_scheduler(schedulers::scheduler_mt::make());
    fg->validate();
    fg->run();
    EXPECT_EQ(snk->data().size(), exp.size());
    EXPECT_TRUE(rel_rms(snk->data(), exp) < TOL);
}

// config 2 as written: device null_source (finite) -> fft -> complex_to_mag -> null_sink, nothing on the host
QA_TEST(Config2, NullSourceToNullSinkOnDevice)
{
    const size_t N = 4096, nv = 4096; // 128 MiB stream
    std::vector<float> w(N);
    orc_window_blackmanharris(w.data(), (int)N);
    auto src = cuda::null_source::make(N * sizeof(gr_complex), nv);
    auto f = cuda::fft::make(N, true, w);
    auto m = cuda::complex_to_mag::make(N);
    auto snk = blocks::null_sink::make(N * sizeof(float));
    auto fg = flowgraph::make();
    fg->connect(src, 0, f, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_D2D);
    fg->connect(f, 0, m, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_D2D);
    fg->connect(m, 0, snk, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_D2D);
    fg->set_scheduler(schedulers::scheduler_mt::make());
    fg->validate();
    fg->run();
    EXPECT_EQ(snk->n_items(), (uint64_t)nv);
}

// BASELINE config 3: fir_filter_ccf (decim 4) -> multiply_const -> fft, device buffers end to end
QA_TEST(Config3, FirMulFftChain)
{
    const int N = 4096, T = 200, D = 4;
    auto in = noise((size_t)N * D * 40, 8);
    auto taps = rtaps(T, 9);
    gr_complex k(0.5f, -0.25f);
    std::vector<float> w(N);
    orc_window_blackmanharris(w.data(), N);
    std::vector<gr_complex> a(in.size() / D), b(a.size()), exp(a.size());
    orc_fir_ccf_f64((float*)a.data(), (const float*)in.data(), (int64_t)in.size(), taps.data(), T, D, nullptr);
    orc_multiply_const_cc((float*)b.data(), (const float*)a.data(), k.real(), k.imag(), (int64_t)a.size());
    orc_fft_f64((float*)exp.data(), (const float*)b.data(), (int64_t)(b.size() / N), N, 1, w.data(), 0);
    for (int fused = 0; fused < 2; fused++) {
        auto src = blocks::vector_source_c::make(in);
        auto fir = cuda::fir_filter_ccf::make(D, taps);
        auto snk = blocks::vector_sink_c::make(N);
        auto fg = flowgraph::make();
        fg->connect(src, 0, fir, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(H2D, 2u << 20));
        if (!fused) {
            auto mul = cuda::multiply_const_cc::make(k);
            auto f = cuda::fft::make(N, true, w, false, cuda::fft_output_t::COMPLEX, /*stream_input=*/true);
            fg->connect(fir, 0, mul, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2D, 2u << 20));
            fg->connect(mul, 0, f, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2D, 2u << 20));
            fg->connect(f, 0, snk, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_D2H);
        } else {
            fir->set_fused_multiply_const(k); // multiply_const folded into the FIR epilogue
            auto f = cuda::fft::make(N, true, w, false, cuda::fft_output_t::COMPLEX, true);
            fg->connect(fir, 0, f, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_D2D);
            fg->connect(f, 0, snk, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_D2H);
        }
        fg->set_scheduler(schedulers::scheduler_mt::make());
        fg->validate();
        fg->run();
        EXPECT_EQ(snk->data().size(), exp.size());
        EXPECT_TRUE(rel_rms(snk->data(), exp) < TOL);
    }
}

// the graph-rewriting fusion pass: fir -> multiply_const -> fft -> complex_to_mag collapses to
// two blocks (FIR with fused k, FFT with fused |.|) and still matches the oracle of the 4-block chain
QA_TEST(Fusion, AdjacentBlocksCollapse)
{
    const int N = 4096, T = 200, D = 4;
    auto in = noise((size_t)N * D * 24, 31);
    auto taps = rtaps(T, 32);
    gr_complex k(0.5f, -0.25f);
    std::vector<float> w(N);
    orc_window_blackmanharris(w.data(), N);
    std::vector<gr_complex> a(in.size() / D), b(a.size()), X(a.size());
    orc_fir_ccf_f64((float*)a.data(), (const float*)in.data(), (int64_t)in.size(), taps.data(), T, D, nullptr);
    orc_multiply_const_cc((float*)b.data(), (const float*)a.data(), k.real(), k.imag(), (int64_t)a.size());
    orc_fft_f64((float*)X.data(), (const float*)b.data(), (int64_t)(b.size() / N), N, 1, w.data(), 0);
    std::vector<float> exp(X.size());
    for (size_t i = 0; i < X.size(); i++)
        exp[i] = (float)std::abs(std::complex<double>(X[i]));

    auto src = blocks::vector_source_c::make(in);
    auto fir = cuda::fir_filter_ccf::make(D, taps);
    auto mul = cuda::multiply_const_cc::make(k);
    auto f = cuda::fft::make(N, true, w, false, cuda::fft_output_t::COMPLEX, /*stream_input=*/true);
    auto mag = cuda::complex_to_mag::make(N);
    auto snk = blocks::vector_sink_f::make(N);
    auto fg = flowgraph::make();
    fg->connect(src, 0, fir, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_H2D);
    fg->connect(fir, 0, mul, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_D2D);
    fg->connect(mul, 0, f, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_D2D);
    fg->connect(f, 0, mag, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_D2D);
    fg->connect(mag, 0, snk, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_D2H);
    EXPECT_EQ(fg->calc_used_blocks().size(), (size_t)6);
    int n = cuda::fuse_adjacent(*fg);
    EXPECT_EQ(n, 2);                                       // fir+mul, fft+mag
    EXPECT_EQ(fg->calc_used_blocks().size(), (size_t)4);   // src, fir, fft, sink
    EXPECT_EQ(fg->edges().size(), (size_t)3);
    for (auto& e : fg->edges())
        EXPECT_TRUE(e->has_custom_buffer());               // device edges survive the rewrite
    fg->set_scheduler(schedulers::scheduler_mt::make());
    fg->validate();
    fg->run();
    EXPECT_EQ(snk->data().size(), exp.size());
    EXPECT_TRUE(rel_rms(snk->data(), exp) < TOL);

    // a fanned-out intermediate stream must NOT be fused away
    auto src2 = blocks::vector_source_c::make(in, false, N);
    auto f2 = cuda::fft::make(N, true, w);
    auto mag2 = cuda::complex_to_mag::make(N);
    auto tap2 = blocks::vector_sink_c::make(N);
    auto snk2 = blocks::vector_sink_f::make(N);
    auto fg2 = flowgraph::make();
    fg2->connect(src2, 0, f2, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_H2D);
    fg2->connect(f2, 0, mag2, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_D2D);
    fg2->connect(f2, 0, tap2, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_D2H);
    fg2->connect(mag2, 0, snk2, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_D2H);
    EXPECT_EQ(cuda::fuse_adjacent(*fg2), 0);
}

// BASELINE config 4 (single GPU slice): 64-channel polyphase channelizer in a flowgraph
QA_TEST(Config4, PfbChannelizer64)
{
    const int M = 64, P = 16;
    auto in = noise((size_t)M * 5000, 10);
    std::vector<float> taps(M * P);
    for (int i = 0; i < M * P; i++) { // windowed sinc prototype
        double t = (i - (M * P - 1) / 2.0) / M;
        double s = std::fabs(t) < 1e-12 ? 1.0 : std::sin(M_PI * t) / (M_PI * t);
        taps[i] = (float)(s * (0.54 - 0.46 * std::cos(2 * M_PI * i / (M * P - 1))) / M);
    }
    std::vector<gr_complex> exp(in.size());
    orc_pfb_channelizer_f64((float*)exp.data(), (const float*)in.data(), (int64_t)in.size(), taps.data(), M, P, nullptr);
    auto src = blocks::vector_source_c::make(in);
    auto ch = cuda::pfb_channelizer_ccf::make(M, taps);
    auto snk = blocks::vector_sink_c::make(M);
    auto fg = flowgraph::make();
    fg->connect(src, 0, ch, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(H2D, 1u << 20));
    fg->connect(ch, 0, snk, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2H, 1u << 20));
    fg->set_scheduler(schedulers::scheduler_mt::make());
    fg->validate();
    fg->run();
    EXPECT_EQ(snk->data().size(), exp.size());
    EXPECT_TRUE(rel_rms(snk->data(), exp) < TOL);
}

// config 4 as BASELINE.json words it ("filterbank + DFT as tensor-core GEMM"): the same flowgraph with the DFT across
// branches on tcgen05 (algorithm 2); small rings, so the stream crosses many work() calls and 64-frame tiles
QA_TEST(Config4, PfbChannelizer64TensorCoreDft)
{
    const int M = 64, P = 16;
    auto in = noise((size_t)M * 5000, 11);
    std::vector<float> taps(M * P);
    for (int i = 0; i < M * P; i++) {
        double t = (i - (M * P - 1) / 2.0) / M;
        double s = std::fabs(t) < 1e-12 ? 1.0 : std::sin(M_PI * t) / (M_PI * t);
        taps[i] = (float)(s * (0.54 - 0.46 * std::cos(2 * M_PI * i / (M * P - 1))) / M);
    }
    std::vector<gr_complex> exp(in.size());
    orc_pfb_channelizer_f64((float*)exp.data(), (const float*)in.data(), (int64_t)in.size(), taps.data(), M, P, nullptr);
    auto src = blocks::vector_source_c::make(in);
    auto ch = cuda::pfb_channelizer_ccf::make(M, taps, 0, 0, 2);
    auto snk = blocks::vector_sink_c::make(M);
    auto fg = flowgraph::make();
    fg->connect(src, 0, ch, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(H2D, 1u << 20));
    fg->connect(ch, 0, snk, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2H, 1u << 20));
    fg->set_scheduler(schedulers::scheduler_mt::make());
    fg->validate();
    fg->run();
    EXPECT_EQ(snk->data().size(), exp.size());
    EXPECT_TRUE(rel_rms(snk->data(), exp) < TOL);
    // the tensor-core form needs 64 channels: anything else must fail loudly, not fall back
    bool threw = false;
    try {
        cuda::pfb_channelizer_ccf::make(32, std::vector<float>(32 * 8, 0.1f), 0, 0, 2);
    } catch (const std::exception&) {
        threw = true;
    }
    EXPECT_TRUE(threw);
}

// two-input sync blocks: both inputs are clamped to the common minimum by sync_block::do_work
QA_TEST(TwoInput, MultiplyAndAdd)
{
    auto a = noise(777777, 21), b = noise(777777, 22);
    auto sa = blocks::vector_source_c::make(a);
    auto sb = blocks::vector_source_c::make(b);
    auto mul = cuda::multiply_cc::make();
    auto add = cuda::add_cc::make();
    auto snk_m = blocks::vector_sink_c::make(), snk_a = blocks::vector_sink_c::make();
    auto fg = flowgraph::make();
    fg->connect(sa, 0, mul, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(H2D, 2u << 20));
    fg->connect(sb, 0, mul, 1)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(H2D, 4u << 20));
    fg->connect(sa, 0, add, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(H2D, 2u << 20));
    fg->connect(sb, 0, add, 1)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(H2D, 2u << 20));
    fg->connect(mul, 0, snk_m, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2H, 2u << 20));
    fg->connect(add, 0, snk_a, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2H, 2u << 20));
    fg->set_scheduler(schedulers::scheduler_mt::make());
    fg->validate();
    fg->run();
    std::vector<gr_complex> em(a.size()), ea(a.size());
    orc_multiply_cc((float*)em.data(), (const float*)a.data(), (const float*)b.data(), (int64_t)a.size());
    orc_add_f((float*)ea.data(), (const float*)a.data(), (const float*)b.data(), (int64_t)a.size() * 2);
    EXPECT_EQ(snk_m->data(), em);
    EXPECT_EQ(snk_a->data(), ea);
}

// stream tags survive device-resident edges (the reference's cuda_buffer breaks them: SURVEY.md 2.3)
QA_TEST(SchedulerMTTags, TagsAcrossDeviceBuffers)
{
    std::vector<gr_complex> in(300000, gr_complex(1, 0));
    std::vector<tag_t> tags;
    for (uint64_t off : { 0ull, 77ull, 150000ull, 299999ull })
        tags.emplace_back(off, pmtf::make_string("k"), pmtf::make_int((int64_t)off));
    auto src = blocks::vector_source_c::make(in, false, 1, tags);
    auto c1 = cuda::copy::make(1), c2 = cuda::copy::make(1);
    auto snk = blocks::vector_sink_c::make();
    auto fg = flowgraph::make();
    fg->connect(src, 0, c1, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(H2D, 2u << 20));
    fg->connect(c1, 0, c2, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2D, 2u << 20));
    fg->connect(c2, 0, snk, 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_SIZED(D2H, 2u << 20));
    fg->set_scheduler(schedulers::scheduler_mt::make());
    fg->validate();
    fg->run();
    auto got = snk->tags();
    EXPECT_EQ(got.size(), tags.size());
    bool ok = got.size() == tags.size();
    for (size_t i = 0; ok && i < tags.size(); i++)
        ok &= got[i] == tags[i];
    EXPECT_TRUE(ok);
}

// BASELINE config 5 as ONE newsched flowgraph on several GPUs of one process (SURVEY.md 8e "single process,
// 8 devices"; reference edges are created per flowgraph in one process,
// schedulers/mt/lib/buffer_management.cpp:78-82): a long stream is cut into time segments, GPU g owns
// segment g resident in its memory, filters it with the 4096-tap FIR and needs the 4095 samples that
// precede it -- a peer copy from GPU g-1's segment, ordered on the FIR block's stream.  Every block,
// ring, stream and event lives on its own device; one mt scheduler drives them all.  Oracle-checked.
static void time_segmented_fir(int G, size_t n_all, int T)
{
    auto x = noise(n_all, 77);
    auto taps = rtaps(T, 78);
    const size_t seg = n_all / G;
    std::vector<std::shared_ptr<cuda::vector_source_c>> src(G);
    std::vector<std::shared_ptr<cuda::fir_filter_ccf>> fir(G);
    std::vector<std::shared_ptr<blocks::vector_sink_c>> snk(G);
    auto fg = flowgraph::make();
    for (int g = 0; g < G; g++) {
        EXPECT_EQ(b200_set_device(g), 0);
        std::vector<gr_complex> part(x.begin() + g * seg, x.begin() + (g + 1) * seg);
        src[g] = cuda::vector_source_c::make(part);
        fir[g] = cuda::fir_filter_ccf::make(1, taps);
        snk[g] = blocks::vector_sink_c::make(1, seg);
        EXPECT_EQ(src[g]->device(), g);
        EXPECT_EQ(fir[g]->device(), g);
        fg->connect(src[g], 0, fir[g], 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_ON(D2D, 8u << 20, g));
        fg->connect(fir[g], 0, snk[g], 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_ON(D2H, 8u << 20, g));
    }
    for (int g = 1; g < G; g++) { // halo: tail of the left neighbour's segment, device to device
        EXPECT_EQ(b200_set_device(g), 0);
        EXPECT_EQ(b200_enable_peer_access(g - 1), 0);
        fir[g]->set_history_device(src[g - 1]->device_data() + seg - (T - 1));
    }
    b200_set_device(0);
    fg->set_scheduler(schedulers::scheduler_mt::make("sched", 32768));
    fg->validate();
    fg->run();
    std::vector<gr_complex> got;
    for (int g = 0; g < G; g++) {
        auto d = snk[g]->data();
        EXPECT_EQ(d.size(), seg);
        got.insert(got.end(), d.begin(), d.end());
    }
    std::vector<gr_complex> ref(n_all);
    orc_fir_ccf_f64((float*)ref.data(), (const float*)x.data(), (int64_t)n_all, taps.data(), T, 1, nullptr);
    const double err = rel_rms(got, ref);
    std::printf("  %d GPUs, %zu samples, %d taps (algorithm %d): rel rms %.3g\n", G, n_all, T, fir[0]->algorithm(), err);
    EXPECT_TRUE(err < TOL);
}
QA_TEST(Config5, TimeSegmentedTwoGpus)
{
    int ngpu = 0;
    b200_device_count(&ngpu);
    if (ngpu < 2) {
        std::printf("  skipped: needs 2 GPUs (found %d)\n", ngpu);
        return;
    }
    time_segmented_fir(2, (size_t)1 << 21, 4096);
    time_segmented_fir(2, (size_t)1 << 20, 128); // tensor-core form
    b200_set_device(0);
}
// the same flowgraph shape on ONE device: every piece of the multi-device plumbing except the peer copy
QA_TEST(Config5, TimeSegmentedOneGpuTwoSegments)
{
    const size_t n_all = (size_t)1 << 20;
    const int T = 4096, G = 2;
    auto x = noise(n_all, 79);
    auto taps = rtaps(T, 80);
    const size_t seg = n_all / G;
    std::vector<std::shared_ptr<cuda::vector_source_c>> src(G);
    std::vector<std::shared_ptr<cuda::fir_filter_ccf>> fir(G);
    std::vector<std::shared_ptr<blocks::vector_sink_c>> snk(G);
    auto fg = flowgraph::make();
    for (int g = 0; g < G; g++) {
        std::vector<gr_complex> part(x.begin() + g * seg, x.begin() + (g + 1) * seg);
        src[g] = cuda::vector_source_c::make(part);
        fir[g] = cuda::fir_filter_ccf::make(1, taps);
        snk[g] = blocks::vector_sink_c::make(1, seg);
        fg->connect(src[g], 0, fir[g], 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_ON(D2D, 8u << 20, 0));
        fg->connect(fir[g], 0, snk[g], 0)->set_custom_buffer(DEVICE_BUFFER_ARGS_ON(D2H, 8u << 20, 0));
    }
    fir[1]->set_history_device(src[0]->device_data() + seg - (T - 1));
    fg->set_scheduler(schedulers::scheduler_mt::make("sched", 32768));
    fg->validate();
    fg->run();
    std::vector<gr_complex> got;
    for (int g = 0; g < G; g++) {
        auto d = snk[g]->data();
        got.insert(got.end(), d.begin(), d.end());
    }
    std::vector<gr_complex> ref(n_all);
    orc_fir_ccf_f64((float*)ref.data(), (const float*)x.data(), (int64_t)n_all, taps.data(), T, 1, nullptr);
    EXPECT_EQ(got.size(), n_all);
    EXPECT_TRUE(rel_rms(got, ref) < TOL);
}

int main(int argc, char** argv) { return qa_main(argc, argv); }
