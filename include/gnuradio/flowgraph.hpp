// gnuradio/flowgraph.hpp -- top-level graph: set_scheduler / validate / start / wait / run, the
// call sequence of every reference test and benchmark (runtime/lib/flowgraph.cpp:51-93,
// schedulers/mt/bench/bm_copy.cpp:145-154).
#pragma once
#include <gnuradio/scheduler.hpp>

namespace gr {

class flowgraph : public graph
{
    std::vector<scheduler_sptr> d_schedulers;

public:
    typedef std::shared_ptr<flowgraph> sptr;
    static sptr make() { return std::make_shared<flowgraph>(); }
    flowgraph() { set_alias("flowgraph"); }
    void set_scheduler(scheduler_sptr s) { d_schedulers = { std::move(s) }; }
    void clear_schedulers() { d_schedulers.clear(); }
    void validate()
    {
        if (d_schedulers.size() != 1)
            throw std::runtime_error("flowgraph: exactly one scheduler must be set");
        d_schedulers[0]->initialize(std::static_pointer_cast<graph>(shared_from_this()));
    }
    void start()
    {
        for (auto& s : d_schedulers)
            s->start();
    }
    void stop()
    {
        for (auto& s : d_schedulers)
            s->stop();
    }
    void wait()
    {
        for (auto& s : d_schedulers)
            s->wait();
    }
    void run()
    {
        start();
        wait();
    }
};
typedef flowgraph::sptr flowgraph_sptr;

} // namespace gr
