// gnuradio/schedulers/mt/scheduler_mt.hpp -- thread-per-block(-group) scheduler.
//
// Host harness with the API and the per-iteration contract of the reference `mt` scheduler
// (schedulers/mt/include/gnuradio/schedulers/mt/scheduler_mt.hpp:11-75, lib/scheduler_mt.cpp:24-84,
// lib/thread_wrapper.cpp:77-191, lib/graph_executor.cpp:7-222, lib/buffer_management.cpp:8-148):
//   * one buffer per edge: the edge's custom factory if set, else the scheduler default
//     (vmcirc), num_items = 2 * fixed_buf_size / itemsize (buffer_management.cpp:117);
//   * one std::thread per block group, then per remaining block;
//   * per iteration and block: read_info on every input (blocked if < 1 item), min write_info
//     over all buffers of every output port (blocked if < 1), do_work(), then post_read /
//     fan-out copy_items / post_write and neighbour notification, tag propagation per policy.
// It exists because the reference runtime cannot be compiled in this image; it
// is a caller of the hot path, not part of it.  Two deliberate differences, both called out in
// SURVEY.md 7.3 as defects of the reference:
//   * end of stream is a real drain: WORK_DONE marks the block's output edges writer_done, and
//     a block finishes once its inputs are exhausted and their writers are done (the reference
//     sleeps 100 ms and drops what is in flight, flowgraph_monitor.cpp:27);
//     and in the other direction a finished consumer marks its input edges reader_done so an
//     endless source upstream (null_source -> head) stops;
//   * a work() call that neither consumes nor produces counts as blocked instead of spinning.
#pragma once
#include <gnuradio/scheduler.hpp>
#include <gnuradio/vmcircbuf.hpp>

#include <pthread.h>
#include <sched.h>

#include <atomic>
#include <condition_variable>
#include <exception>
#include <map>
#include <thread>

namespace gr {
namespace schedulers {

class buffer_manager
{
    std::map<edge*, buffer_sptr> _edge_buf;
    std::map<port_base*, buffer_sptr> _in_buf;
    std::map<port_base*, std::vector<buffer_sptr>> _out_bufs;
    int _fixed_buf_size;

public:
    typedef std::shared_ptr<buffer_manager> sptr;
    explicit buffer_manager(int fixed_buf_size) : _fixed_buf_size(fixed_buf_size) {}
    void initialize_buffers(flat_graph_sptr fg, buffer_factory_function def_factory,
                            std::shared_ptr<buffer_properties> def_props)
    {
        for (auto& e : fg->edges()) {
            size_t itemsize = e->itemsize();
            size_t num_items = std::max<size_t>(2, 2 * (size_t)_fixed_buf_size / itemsize);
            buffer_sptr b = e->has_custom_buffer() ? e->buffer_factory()(num_items, itemsize, e->buf_properties())
                                                   : def_factory(num_items, itemsize, def_props);
            b->set_name(e->identifier());
            _edge_buf[e.get()] = b;
            _in_buf[e->dst().port().get()] = b;
            _out_bufs[e->src().port().get()].push_back(b);
        }
    }
    buffer_sptr get_input_buffer(port_sptr p)
    {
        auto it = _in_buf.find(p.get());
        return it == _in_buf.end() ? nullptr : it->second;
    }
    // read-only after initialize_buffers(): several scheduler threads call this concurrently, so no operator[]
    const std::vector<buffer_sptr>& get_output_buffers(port_sptr p) const
    {
        static const std::vector<buffer_sptr> none;
        auto it = _out_bufs.find(p.get());
        return it == _out_bufs.end() ? none : it->second;
    }
};

enum class executor_iteration_status { READY, BLKD_IN, BLKD_OUT, DONE };

class thread_wrapper : public neighbor_interface
{
    std::vector<block_sptr> _blocks;
    std::vector<bool> _finished;
    buffer_manager::sptr _bufman;
    std::thread _thread;
    std::mutex _m;
    std::condition_variable _cv;
    bool _notified = false;
    std::atomic<bool> _stop{ false }, _started{ false };
    std::exception_ptr _error = nullptr;
    std::vector<unsigned int> _affinity; // CPUs this thread may run on (empty: anywhere); thread_wrapper.cpp:23-31
    bool _rt_prio = false;

    void finish_block(size_t bi)
    {
        _finished[bi] = true;
        auto& b = _blocks[bi];
        for (auto& p : b->output_stream_ports())
            for (auto& buf : _bufman->get_output_buffers(p))
                buf->set_writer_done();
        for (auto& p : b->input_stream_ports())
            if (auto buf = _bufman->get_input_buffer(p))
                buf->set_reader_done();
        for (auto& p : b->output_stream_ports())
            p->notify_connected_ports(std::make_shared<scheduler_action>(scheduler_action_t::NOTIFY_INPUT));
        for (auto& p : b->input_stream_ports())
            p->notify_connected_ports(std::make_shared<scheduler_action>(scheduler_action_t::NOTIFY_OUTPUT));
        b->done();
    }

    // one block, one iteration: graph_executor.cpp:17-218
    executor_iteration_status run_block(size_t bi)
    {
        auto& b = _blocks[bi];
        std::vector<block_work_input> work_input;
        std::vector<block_work_output> work_output;
        auto in_ports = b->input_stream_ports();
        auto out_ports = b->output_stream_ports();

        bool blocked = false, exhausted = false, all_upstream_done = !in_ports.empty();
        for (auto& p : in_ports) {
            auto buf = _bufman->get_input_buffer(p);
            if (!buf)
                throw std::runtime_error("unconnected input port on " + b->alias());
            // sample the flag BEFORE read_info: the writer posts its last items, then sets it
            bool wd = buf->writer_done();
            buffer_info_t ri;
            buf->read_info(ri);
            if (ri.n_items < 1) {
                if (wd)
                    exhausted = true; // nothing left and nothing more will come
                else
                    blocked = true;
            }
            if (!wd)
                all_upstream_done = false;
            work_input.emplace_back(ri.n_items, buf);
        }
        if (blocked)
            return executor_iteration_status::BLKD_IN;
        if (exhausted) {
            finish_block(bi);
            return executor_iteration_status::DONE;
        }
        // nobody downstream is listening any more (e.g. a head block finished): stop producing
        for (auto& p : out_ports) {
            auto& bufs = _bufman->get_output_buffers(p);
            bool all_gone = !bufs.empty();
            for (auto& buf : bufs)
                all_gone &= buf->reader_done();
            if (all_gone) {
                finish_block(bi);
                return executor_iteration_status::DONE;
            }
        }
        for (auto& p : out_ports) {
            int space = std::numeric_limits<int>::max();
            buffer_sptr first = nullptr;
            for (auto& buf : _bufman->get_output_buffers(p)) {
                buffer_info_t wi;
                buf->write_info(wi);
                space = std::min(space, wi.n_items);
                if (!first)
                    first = buf;
            }
            if (!first)
                throw std::runtime_error("unconnected output port on " + b->alias());
            if (space < 1)
                return executor_iteration_status::BLKD_OUT;
            work_output.emplace_back(space, first);
        }
        std::vector<int> given_space;
        for (auto& w : work_output)
            given_space.push_back(w.n_items);

        work_return_code_t ret = b->do_work(work_input, work_output);
        if (ret != work_return_code_t::WORK_OK && ret != work_return_code_t::WORK_DONE)
            throw std::runtime_error("block " + b->alias() + " returned an error from work()");

        bool progress = false;
        for (size_t i = 0; i < in_ports.size(); i++) {
            auto& w = work_input[i];
            int nc = std::max(0, w.n_consumed);
            if (w.buffer->has_tags()) {
                auto pol = b->tag_propagation_policy();
                for (size_t o = 0; o < out_ports.size(); o++)
                    if (pol == tag_propagation_policy_t::TPP_ALL_TO_ALL ||
                        (pol == tag_propagation_policy_t::TPP_ONE_TO_ONE && o == i))
                        for (auto& ob : _bufman->get_output_buffers(out_ports[o]))
                            ob->propagate_tags(w.buffer, nc, std::max(0, work_output[o].n_produced));
                w.buffer->prune_tags(nc);
            }
            if (nc > 0) {
                w.buffer->post_read(nc);
                progress = true;
                in_ports[i]->notify_connected_ports(
                    std::make_shared<scheduler_action>(scheduler_action_t::NOTIFY_OUTPUT));
            }
        }
        for (size_t o = 0; o < out_ports.size(); o++) {
            int np = std::max(0, work_output[o].n_produced);
            if (np > 0) {
                auto& bufs = _bufman->get_output_buffers(out_ports[o]);
                for (size_t j = 1; j < bufs.size(); j++)
                    bufs[j]->copy_items(bufs[0], np);
                for (auto& buf : bufs)
                    buf->post_write(np);
                progress = true;
                out_ports[o]->notify_connected_ports(
                    std::make_shared<scheduler_action>(scheduler_action_t::NOTIFY_INPUT));
            }
        }
        if (ret == work_return_code_t::WORK_DONE) {
            finish_block(bi);
            return executor_iteration_status::DONE;
        }
        if (!progress) {
            // 0/0 work: blocked, unless nothing more can ever arrive.  A block with an output
            // multiple (an interpolator needs L free items) may also have produced nothing only
            // because downstream has not drained yet: its space still grows, so that is BLKD_OUT,
            // not the end of the stream.
            if (all_upstream_done) {
                // Order matters (this thread is the only writer of its output rings; the readers can only empty
                // them).  (1) Anything still pending downstream: the space will grow, wait for it.  (2) The rings
                // are empty -- and stay empty -- so the space seen NOW is all there will ever be: if it is more than
                // this call was given (the reader drained the ring after write_info above), call again; only a call
                // that had all of it and still produced nothing proves that nothing more can come out.  Checking
                // the space first and "pending" second let the reader drain in between: the block retired with up
                // to a ring of input unplaced (a 1-in-8 failure of SchedulerMTTest.OutputMultipleDrain).
                bool output_pending = false;
                for (auto& p : out_ports)
                    for (auto& buf : _bufman->get_output_buffers(p))
                        output_pending |= !buf->reader_done() && buf->items_pending() != 0;
                if (output_pending)
                    return executor_iteration_status::BLKD_OUT;
                for (size_t o = 0; o < out_ports.size(); o++) {
                    int space = std::numeric_limits<int>::max();
                    for (auto& buf : _bufman->get_output_buffers(out_ports[o])) {
                        buffer_info_t wi;
                        buf->write_info(wi);
                        space = std::min(space, wi.n_items);
                    }
                    if (space > given_space[o])
                        return executor_iteration_status::READY;
                }
                finish_block(bi);
                return executor_iteration_status::DONE;
            }
            return executor_iteration_status::BLKD_IN;
        }
        return executor_iteration_status::READY;
    }

    void thread_body()
    {
        try {
            thread_loop();
        } catch (...) {
            // a block threw (e.g. a CUDA error surfaced through the C-ABI): remember it for wait()
            // and release everybody else instead of std::terminate (the reference has no handler,
            // SURVEY.md section 5 "Failure detection")
            _error = std::current_exception();
            for (size_t bi = 0; bi < _blocks.size(); bi++)
                if (!_finished[bi]) {
                    try {
                        finish_block(bi);
                    } catch (...) {
                    }
                }
        }
    }

    void thread_loop()
    {
        while (!_stop.load()) {
            bool any_ready = false, all_finished = true;
            for (size_t bi = 0; bi < _blocks.size(); bi++) {
                if (_finished[bi])
                    continue;
                auto st = run_block(bi);
                if (st == executor_iteration_status::READY)
                    any_ready = true;
                if (!_finished[bi])
                    all_finished = false;
            }
            if (all_finished)
                break;
            if (!any_ready) {
                std::unique_lock<std::mutex> lk(_m);
                _cv.wait_for(lk, std::chrono::milliseconds(2), [this] { return _notified || _stop.load(); });
                _notified = false;
            }
        }
    }

public:
    typedef std::shared_ptr<thread_wrapper> sptr;
    thread_wrapper(std::vector<block_sptr> blocks, buffer_manager::sptr bufman)
        : _blocks(std::move(blocks)), _finished(_blocks.size(), false), _bufman(std::move(bufman))
    {
    }
    void push_message(scheduler_message_sptr) override
    {
        {
            std::lock_guard<std::mutex> lk(_m);
            _notified = true;
        }
        _cv.notify_one();
    }
    void set_affinity(const std::vector<unsigned int>& cpus) { _affinity = cpus; }
    void set_rt_prio(bool on) { _rt_prio = on; }
    void start()
    {
        _started = true;
        _thread = std::thread([this] { thread_body(); });
        if (!_affinity.empty()) {
            cpu_set_t set;
            CPU_ZERO(&set);
            for (auto c : _affinity)
                CPU_SET(c % CPU_SETSIZE, &set);
            pthread_setaffinity_np(_thread.native_handle(), sizeof(set), &set); // best effort, like the reference
        }
        if (_rt_prio) {
            sched_param sp{};
            sp.sched_priority = sched_get_priority_min(SCHED_FIFO);
            pthread_setschedparam(_thread.native_handle(), SCHED_FIFO, &sp); // needs CAP_SYS_NICE; ignored otherwise
        }
    }
    void stop()
    {
        _stop = true;
        push_message(nullptr);
        for (auto& b : _blocks)
            b->stop();
    }
    void wait()
    {
        if (_thread.joinable())
            _thread.join();
    }
    std::exception_ptr error() const { return _error; }
};

class scheduler_mt : public scheduler
{
    const int s_fixed_buf_size;
    std::vector<std::vector<block_sptr>> _block_groups;
    std::vector<std::vector<unsigned int>> _group_affinity;
    std::vector<unsigned int> _thread_cpus; // round-robin over the per-block threads (bm_copy.cpp --cpus)
    bool _rt_prio = false;
    std::vector<thread_wrapper::sptr> _threads;
    buffer_manager::sptr _bufman;

public:
    typedef std::shared_ptr<scheduler_mt> sptr;
    static sptr make(const std::string name = "multi_threaded", const unsigned int fixed_buf_size = 32768)
    {
        return std::make_shared<scheduler_mt>(name, fixed_buf_size);
    }
    scheduler_mt(const std::string name = "multi_threaded", const unsigned int fixed_buf_size = 32768)
        : scheduler(name), s_fixed_buf_size((int)fixed_buf_size)
    {
        _default_buf_factory = vmcirc_buffer::make;
        _default_buf_properties = vmcirc_buffer_properties::make(vmcirc_buffer_type::AUTO);
    }
    void add_block_group(const std::vector<block_sptr>& blocks, const std::string& = "",
                         const std::vector<unsigned int>& cpu_affinity = {})
    {
        _block_groups.push_back(blocks);
        _group_affinity.push_back(cpu_affinity);
    }
    void set_thread_affinity(const std::vector<unsigned int>& cpus) { _thread_cpus = cpus; }
    void set_rt_prio(bool on) { _rt_prio = on; }
    buffer_manager::sptr buffers() { return _bufman; }

    void initialize(flat_graph_sptr fg) override
    {
        _bufman = std::make_shared<buffer_manager>(s_fixed_buf_size);
        _bufman->initialize_buffers(fg, _default_buf_factory, _default_buf_properties);
        auto blocks = fg->calc_used_blocks();
        std::vector<block_sptr> grouped;
        auto make_thread = [&](const std::vector<block_sptr>& grp, const std::vector<unsigned int>& aff) {
            auto t = std::make_shared<thread_wrapper>(grp, _bufman);
            for (auto& b : grp)
                for (auto& p : b->all_ports())
                    p->set_parent_intf(t);
            if (!aff.empty())
                t->set_affinity(aff);
            else if (!_thread_cpus.empty())
                t->set_affinity({ _thread_cpus[_threads.size() % _thread_cpus.size()] });
            t->set_rt_prio(_rt_prio);
            _threads.push_back(t);
        };
        for (size_t gi = 0; gi < _block_groups.size(); gi++) {
            auto& grp = _block_groups[gi];
            make_thread(grp, _group_affinity[gi]);
            grouped.insert(grouped.end(), grp.begin(), grp.end());
        }
        for (auto& b : blocks)
            if (std::find(grouped.begin(), grouped.end(), b) == grouped.end())
                make_thread({ b }, {});
        for (auto& b : blocks) {
            std::vector<buffer_sptr> ins;
            std::vector<std::vector<buffer_sptr>> outs;
            for (auto& p : b->input_stream_ports())
                ins.push_back(_bufman->get_input_buffer(p));
            for (auto& p : b->output_stream_ports())
                outs.push_back(_bufman->get_output_buffers(p));
            b->buffers_attached(ins, outs);
        }
    }
    void start() override
    {
        for (auto& t : _threads)
            t->start();
    }
    void stop() override
    {
        for (auto& t : _threads)
            t->stop();
    }
    void wait() override
    {
        for (auto& t : _threads)
            t->wait();
        for (auto& t : _threads)
            if (t->error())
                std::rethrow_exception(t->error()); // first failure, after every thread has stopped
    }
    // Drops the threads and the edge buffers (device rings, pinned staging, streams, events).  Not done inside
    // wait(): unmapping a 256 MiB VMM ring or unpinning 128 MiB of host memory takes milliseconds and would
    // land inside every start()->wait() timing.  The destructor calls it; nothing keeps the scheduler alive
    // once the flowgraph goes (ports and neighbours hold weak references only).
    void release()
    {
        for (auto& t : _threads) {
            t->stop();
            t->wait();
        }
        _threads.clear();
        _bufman.reset();
    }
    ~scheduler_mt() override { release(); }
};

} // namespace schedulers
} // namespace gr
