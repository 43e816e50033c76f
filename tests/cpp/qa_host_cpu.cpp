// CPU-only QA of the host boundary (block / buffer / edge / flowgraph / mt scheduler harness):
// mirrors the shape of the reference's scheduler tests (schedulers/mt/test/qa_scheduler_mt.cpp,
// qa_block_grouping.cpp, qa_tags.cpp) with host blocks only, so it runs without a GPU.
#include <gnuradio/blocklib/blocks/head.hpp>
#include <gnuradio/blocklib/blocks/null_sink.hpp>
#include <gnuradio/blocklib/blocks/null_source.hpp>
#include <gnuradio/blocklib/blocks/vector_sink.hpp>
#include <gnuradio/blocklib/blocks/vector_source.hpp>
#include <gnuradio/flowgraph.hpp>
#include <gnuradio/schedulers/mt/scheduler_mt.hpp>

#include "qa_common.hpp"

using namespace gr;

// test-only CPU blocks (the product has no CPU signal path): y = k*x and decimate-by-D
class cpu_scale : public sync_block
{
    gr_complex k;

public:
    static std::shared_ptr<cpu_scale> make(gr_complex k_)
    {
        auto p = std::make_shared<cpu_scale>(k_);
        p->add_port(port<gr_complex>::make("input", port_direction_t::INPUT));
        p->add_port(port<gr_complex>::make("output", port_direction_t::OUTPUT));
        return p;
    }
    explicit cpu_scale(gr_complex k_) : sync_block("cpu_scale"), k(k_) {}
    work_return_code_t work(std::vector<block_work_input>& in, std::vector<block_work_output>& out) override
    {
        auto* i = (const gr_complex*)in[0].buffer->read_ptr();
        auto* o = (gr_complex*)out[0].buffer->write_ptr();
        for (int n = 0; n < out[0].n_items; n++)
            o[n] = i[n] * k;
        out[0].n_produced = out[0].n_items;
        return work_return_code_t::WORK_OK;
    }
};

class cpu_keep_one_in_n : public block
{
    int D;

public:
    static std::shared_ptr<cpu_keep_one_in_n> make(int D_)
    {
        auto p = std::make_shared<cpu_keep_one_in_n>(D_);
        p->add_port(port<float>::make("input", port_direction_t::INPUT));
        p->add_port(port<float>::make("output", port_direction_t::OUTPUT));
        return p;
    }
    explicit cpu_keep_one_in_n(int D_) : block("keep_one_in_n"), D(D_) {}
    work_return_code_t work(std::vector<block_work_input>& in, std::vector<block_work_output>& out) override
    {
        int n = std::min(in[0].n_items / D, out[0].n_items);
        auto* i = (const float*)in[0].buffer->read_ptr();
        auto* o = (float*)out[0].buffer->write_ptr();
        for (int m = 0; m < n; m++)
            o[m] = i[m * D];
        in[0].n_consumed = n * D;
        out[0].n_produced = n;
        return work_return_code_t::WORK_OK;
    }
};

// qa_scheduler_mt.cpp:17-39
QA_TEST(SchedulerMTTest, TwoSinks)
{
    std::vector<float> input_data{ 1.0, 2.0, 3.0, 4.0, 5.0 };
    auto src = blocks::vector_source_f::make(input_data, false);
    auto snk1 = blocks::vector_sink_f::make();
    auto snk2 = blocks::vector_sink_f::make();
    auto fg = flowgraph::make();
    fg->connect(src, 0, snk1, 0);
    fg->connect(src, 0, snk2, 0);
    auto sched = schedulers::scheduler_mt::make();
    fg->set_scheduler(sched);
    fg->validate();
    fg->start();
    fg->wait();
    EXPECT_EQ(snk1->data(), input_data);
    EXPECT_EQ(snk2->data(), input_data);
}

// qa_scheduler_mt.cpp:79-135 (shape): 1e6-sample ramp fanned out to N x [scale(k=1) -> sink]
QA_TEST(SchedulerMTTest, BlockFanout)
{
    int num_samples = 1000000;
    std::vector<gr_complex> input_data(num_samples);
    for (int i = 0; i < num_samples; i++)
        input_data[i] = gr_complex(2 * i, 2 * i + 1);
    for (auto nblocks : { 2, 8, 16 }) {
        auto src = blocks::vector_source_c::make(input_data);
        std::vector<std::shared_ptr<blocks::vector_sink_c>> sinks(nblocks);
        auto fg = flowgraph::make();
        for (int i = 0; i < nblocks; i++) {
            auto mult = cpu_scale::make(gr_complex(1.0f, 0.0f));
            sinks[i] = blocks::vector_sink_c::make();
            fg->connect(src, 0, mult, 0);
            fg->connect(mult, 0, sinks[i], 0);
        }
        auto sched = schedulers::scheduler_mt::make("sched", 8192);
        fg->set_scheduler(sched);
        fg->validate();
        fg->start();
        fg->wait();
        for (int n = 0; n < nblocks; n++) {
            EXPECT_EQ(sinks[n]->data().size(), input_data.size()); // nothing dropped at end of stream
            EXPECT_EQ(sinks[n]->data(), input_data);
        }
    }
}

// qa_block_grouping.cpp:15-66 (shape): chains split into block groups
QA_TEST(SchedulerBlockGrouping, BasicBlockGrouping)
{
    int num_samples = 200000;
    std::vector<gr_complex> input_data(num_samples);
    for (int i = 0; i < num_samples; i++)
        input_data[i] = gr_complex(2 * i, 2 * i + 1);
    for (auto ngroups : { 2, 4 })
        for (auto nblocks : { 2, 8 }) {
            auto src = blocks::vector_source_c::make(input_data);
            auto snk = blocks::vector_sink_c::make();
            auto fg = flowgraph::make();
            auto sched = schedulers::scheduler_mt::make("sched", 32768);
            std::vector<std::vector<block_sptr>> groups(ngroups);
            node_sptr last = src;
            for (int g = 0; g < ngroups; g++)
                for (int b = 0; b < nblocks; b++) {
                    auto m = cpu_scale::make(gr_complex(1.0f, 0.0f));
                    fg->connect(last, 0, m, 0);
                    last = m;
                    groups[g].push_back(m);
                }
            fg->connect(last, 0, snk, 0);
            for (auto& g : groups)
                sched->add_block_group(g);
            fg->set_scheduler(sched);
            fg->validate();
            fg->run();
            EXPECT_EQ(snk->data(), input_data);
        }
}

QA_TEST(SchedulerMTTest, NullSourceHeadCounts)
{
    auto src = blocks::null_source::make(sizeof(gr_complex));
    auto hd = blocks::head::make(sizeof(gr_complex), 40000);
    auto snk = blocks::vector_sink_c::make();
    auto fg = flowgraph::make();
    fg->connect(src, 0, hd, 0);
    fg->connect(hd, 0, snk, 0);
    fg->set_scheduler(schedulers::scheduler_mt::make());
    fg->validate();
    fg->run();
    EXPECT_EQ(snk->data().size(), (size_t)40000);
    bool all_zero = true;
    for (auto& v : snk->data())
        all_zero &= (v == gr_complex(0, 0));
    EXPECT_TRUE(all_zero);
}

// test-only interpolator: every input item repeated L times; needs L free output items per input,
// so it returns 0/0 whenever downstream has left it fewer than L
class cpu_repeat : public block
{
    int L;

public:
    static std::shared_ptr<cpu_repeat> make(int L_)
    {
        auto p = std::make_shared<cpu_repeat>(L_);
        p->add_port(port<float>::make("input", port_direction_t::INPUT));
        p->add_port(port<float>::make("output", port_direction_t::OUTPUT));
        return p;
    }
    explicit cpu_repeat(int L_) : block("repeat"), L(L_) {}
    work_return_code_t work(std::vector<block_work_input>& in, std::vector<block_work_output>& out) override
    {
        int n = std::min(in[0].n_items, out[0].n_items / L);
        auto* i = (const float*)in[0].buffer->read_ptr();
        auto* o = (float*)out[0].buffer->write_ptr();
        for (int m = 0; m < n; m++)
            for (int r = 0; r < L; r++)
                o[m * L + r] = i[m];
        in[0].n_consumed = n;
        out[0].n_produced = n * L;
        return work_return_code_t::WORK_OK;
    }
};

// A block with an output multiple must not be retired at end of stream just because downstream
// momentarily left it fewer than L free items (0/0 work with all inputs done): every input item
// has to come out L times.  Small rings make the 0/0-while-output-pending case certain.
QA_TEST(SchedulerMTTest, OutputMultipleDrain)
{
    for (int L : { 2, 5, 64, 1000 }) {
        std::vector<float> in(20011);
        for (size_t i = 0; i < in.size(); i++)
            in[i] = (float)i;
        auto src = blocks::vector_source_f::make(in);
        auto rep = cpu_repeat::make(L);
        auto snk = blocks::vector_sink_f::make();
        auto fg = flowgraph::make();
        fg->connect(src, 0, rep, 0);
        fg->connect(rep, 0, snk, 0);
        fg->set_scheduler(schedulers::scheduler_mt::make("s", 8192));
        fg->validate();
        fg->run();
        auto out = snk->data();
        EXPECT_EQ(out.size(), in.size() * (size_t)L);
        bool ok = out.size() == in.size() * (size_t)L;
        for (size_t m = 0; ok && m < out.size(); m++)
            ok &= out[m] == in[m / L];
        EXPECT_TRUE(ok);
    }
}

// rate-changing gr::block (n_consumed != n_produced), 0/0 work calls and the tail shorter than D
QA_TEST(SchedulerMTTest, RateChangeAndDrain)
{
    for (int D : { 1, 3, 7, 4096 }) {
        std::vector<float> in(100003);
        for (size_t i = 0; i < in.size(); i++)
            in[i] = (float)i;
        auto src = blocks::vector_source_f::make(in);
        auto dec = cpu_keep_one_in_n::make(D);
        auto snk = blocks::vector_sink_f::make();
        auto fg = flowgraph::make();
        fg->connect(src, 0, dec, 0);
        fg->connect(dec, 0, snk, 0);
        fg->set_scheduler(schedulers::scheduler_mt::make("s", 65536));
        fg->validate();
        fg->run();
        auto out = snk->data();
        EXPECT_EQ(out.size(), in.size() / D);
        bool ok = true;
        for (size_t m = 0; m < out.size(); m++)
            ok &= out[m] == in[m * D];
        EXPECT_TRUE(ok);
    }
}

// qa_tags.cpp (shape): tags ride along, offsets stay absolute, every policy
QA_TEST(SchedulerMTTags, PropagationPolicies)
{
    std::vector<gr_complex> in(50000, gr_complex(1, 0));
    std::vector<tag_t> tags;
    for (uint64_t off : { 0ull, 1ull, 4999ull, 25000ull, 49999ull })
        tags.emplace_back(off, pmtf::make_string("k"), pmtf::make_int((int64_t)off));
    for (auto pol : { tag_propagation_policy_t::TPP_ALL_TO_ALL, tag_propagation_policy_t::TPP_ONE_TO_ONE,
                      tag_propagation_policy_t::TPP_DONT }) {
        auto src = blocks::vector_source_c::make(in, false, 1, tags);
        auto a = cpu_scale::make(gr_complex(1, 0));
        auto b = cpu_scale::make(gr_complex(1, 0));
        a->set_tag_propagation_policy(pol);
        b->set_tag_propagation_policy(pol);
        auto snk = blocks::vector_sink_c::make();
        auto fg = flowgraph::make();
        fg->connect(src, 0, a, 0);
        fg->connect(a, 0, b, 0);
        fg->connect(b, 0, snk, 0);
        fg->set_scheduler(schedulers::scheduler_mt::make("s", 4096));
        fg->validate();
        fg->run();
        EXPECT_EQ(snk->data().size(), in.size());
        auto got = snk->tags();
        if (pol == tag_propagation_policy_t::TPP_DONT) {
            EXPECT_EQ(got.size(), (size_t)0);
        } else {
            EXPECT_EQ(got.size(), tags.size());
            bool ok = got.size() == tags.size();
            for (size_t i = 0; ok && i < tags.size(); i++)
                ok &= got[i] == tags[i];
            EXPECT_TRUE(ok);
        }
    }
}

// tags through a rate-changing block land on the decimated offsets
QA_TEST(SchedulerMTTags, DecimationScalesOffsets)
{
    const int D = 4;
    std::vector<float> in(40000);
    for (size_t i = 0; i < in.size(); i++)
        in[i] = (float)i;
    std::vector<tag_t> tags;
    for (uint64_t off : { 0ull, 8ull, 4000ull, 39996ull })
        tags.emplace_back(off, pmtf::make_string("k"), pmtf::make_int((int64_t)off));
    auto src = blocks::vector_source_f::make(in, false, 1, tags);
    auto dec = cpu_keep_one_in_n::make(D);
    auto snk = blocks::vector_sink_f::make();
    auto fg = flowgraph::make();
    fg->connect(src, 0, dec, 0);
    fg->connect(dec, 0, snk, 0);
    fg->set_scheduler(schedulers::scheduler_mt::make("s", 4096));
    fg->validate();
    fg->run();
    auto got = snk->tags();
    EXPECT_EQ(got.size(), tags.size());
    bool ok = got.size() == tags.size();
    for (size_t i = 0; ok && i < tags.size(); i++)
        ok &= got[i].offset == tags[i].offset / D && pmtf::equal(got[i].value, tags[i].value);
    EXPECT_TRUE(ok);
}

// a throwing block must end the flowgraph and surface in wait(), not abort the process
class cpu_thrower : public sync_block
{
    int calls = 0;

public:
    static std::shared_ptr<cpu_thrower> make()
    {
        auto p = std::make_shared<cpu_thrower>();
        p->add_port(port<float>::make("input", port_direction_t::INPUT));
        p->add_port(port<float>::make("output", port_direction_t::OUTPUT));
        return p;
    }
    cpu_thrower() : sync_block("thrower") {}
    work_return_code_t work(std::vector<block_work_input>& in, std::vector<block_work_output>& out) override
    {
        if (++calls > 3)
            throw std::runtime_error("device fell off the bus");
        memcpy(out[0].buffer->write_ptr(), in[0].buffer->read_ptr(), sizeof(float) * out[0].n_items);
        out[0].n_produced = out[0].n_items;
        return work_return_code_t::WORK_OK;
    }
};

QA_TEST(SchedulerMTTest, BlockExceptionSurfacesInWait)
{
    auto src = blocks::null_source::make(sizeof(float)); // endless
    auto thr = cpu_thrower::make();
    auto snk = blocks::null_sink::make(sizeof(float));
    auto fg = flowgraph::make();
    fg->connect(src, 0, thr, 0);
    fg->connect(thr, 0, snk, 0);
    fg->set_scheduler(schedulers::scheduler_mt::make());
    fg->validate();
    fg->start();
    bool caught = false;
    try {
        fg->wait();
    } catch (const std::runtime_error& e) {
        caught = std::string(e.what()).find("fell off") != std::string::npos;
    }
    EXPECT_TRUE(caught);
}

// thread_wrapper -> block -> port -> thread_wrapper and port <-> port were shared_ptr
// cycles (as in the reference), so a flowgraph -- and every edge buffer: 64 MiB device rings plus pinned
// staging here -- was never freed.  A counting buffer type shows the edges are destroyed when the run
// is over, and the blocks when the last outside reference goes away.
static std::atomic<int> g_live_buffers{ 0 };
class counted_buffer : public vmcirc_buffer
{
public:
    counted_buffer(size_t n, size_t isz) : vmcirc_buffer(n, isz) { g_live_buffers++; }
    ~counted_buffer() override { g_live_buffers--; }
    static buffer_sptr make(size_t n, size_t isz, std::shared_ptr<buffer_properties>)
    {
        return buffer_sptr(new counted_buffer(n, isz));
    }
};
QA_TEST(SchedulerMTTest, TeardownFreesBuffersAndBlocks)
{
    std::weak_ptr<blocks::vector_sink_f> weak_sink;
    std::weak_ptr<flowgraph> weak_fg;
    for (int round = 0; round < 3; round++) {
        std::vector<float> input_data(50000, 1.0f);
        auto src = blocks::vector_source_f::make(input_data, false);
        auto hd = blocks::head::make(sizeof(float), 40000);
        auto snk = blocks::vector_sink_f::make();
        auto fg = flowgraph::make();
        fg->connect(src, 0, hd, 0)->set_custom_buffer(counted_buffer::make, nullptr);
        fg->connect(hd, 0, snk, 0)->set_custom_buffer(counted_buffer::make, nullptr);
        auto sched = schedulers::scheduler_mt::make();
        fg->set_scheduler(sched);
        fg->validate();
        EXPECT_EQ(g_live_buffers.load(), 2);
        fg->start();
        fg->wait();
        EXPECT_EQ(snk->data().size(), (size_t)40000);
        if (round == 1) {
            sched->release(); // explicit, e.g. before building the next flowgraph in the same scope
            EXPECT_EQ(g_live_buffers.load(), 0);
        }
        weak_sink = snk;
        weak_fg = fg;
    }
    EXPECT_EQ(g_live_buffers.load(), 0); // edge buffers go with the flowgraph, not with the process
    EXPECT_TRUE(weak_sink.expired());
    EXPECT_TRUE(weak_fg.expired());
}

QA_TEST(Buffers, VmcircWindowIsLinear)
{
    auto buf = vmcirc_buffer::make(1024, sizeof(int), nullptr);
    buffer_info_t wi, ri;
    int next = 0, expect = 0;
    for (int round = 0; round < 200; round++) {
        buf->write_info(wi);
        int n = std::min(wi.n_items, 300);
        for (int i = 0; i < n; i++)
            ((int*)wi.ptr)[i] = next++;
        buf->post_write(n);
        buf->read_info(ri);
        bool ok = true;
        for (int i = 0; i < ri.n_items; i++)
            ok &= ((int*)ri.ptr)[i] == expect++;
        EXPECT_TRUE(ok);
        buf->post_read(ri.n_items);
    }
    EXPECT_EQ(buf->total_read(), buf->total_written());
}

QA_TEST(Graph, ConnectRules)
{
    auto src = blocks::vector_source_f::make({ 1.f });
    auto snk = blocks::vector_sink_f::make();
    auto fg = flowgraph::make();
    auto e = fg->connect(src, 0, snk, 0);
    EXPECT_EQ(e->itemsize(), sizeof(float));
    EXPECT_TRUE(!e->has_custom_buffer());
    e->set_custom_buffer(vmcirc_buffer::make, vmcirc_buffer_properties::make());
    EXPECT_TRUE(e->has_custom_buffer());
    bool threw = false;
    try {
        fg->connect(src, 0, snk, 0); // input already connected
    } catch (const std::invalid_argument&) {
        threw = true;
    }
    EXPECT_TRUE(threw);
    threw = false;
    try {
        fg->connect(src, 3, snk, 0);
    } catch (const std::invalid_argument&) {
        threw = true;
    }
    EXPECT_TRUE(threw);
}

int main(int argc, char** argv) { return qa_main(argc, argv); }
