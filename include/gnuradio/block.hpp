// gnuradio/block.hpp -- abstract signal processing block.
// The virtual interface the `mt` scheduler calls, as in reference
// runtime/include/gnuradio/block.hpp:24-104: start/stop/done (:45-60), work (:81-85, throws if not
// overridden), do_work (:95-99).
#pragma once
#include <gnuradio/block_work_io.hpp>
#include <gnuradio/node.hpp>

#include <stdexcept>

namespace gr {

class scheduler;

class block : public node, public std::enable_shared_from_this<block>
{
private:
    bool d_running = false;
    tag_propagation_policy_t d_tag_propagation_policy = tag_propagation_policy_t::TPP_ALL_TO_ALL;

protected:
    std::shared_ptr<scheduler> p_scheduler = nullptr;

public:
    explicit block(const std::string& name) : node(name) {}
    virtual ~block() {}

    virtual bool start() { d_running = true; return true; }
    virtual bool stop() { d_running = false; return true; }
    virtual bool done() { d_running = false; return true; }

    typedef std::shared_ptr<block> sptr;
    sptr base() { return shared_from_this(); }

    tag_propagation_policy_t tag_propagation_policy() { return d_tag_propagation_policy; }
    void set_tag_propagation_policy(tag_propagation_policy_t p) { d_tag_propagation_policy = p; }

    virtual work_return_code_t work(std::vector<block_work_input>&, std::vector<block_work_output>&)
    {
        throw std::runtime_error("work function has been called but not implemented");
    }
    virtual work_return_code_t do_work(std::vector<block_work_input>& work_input,
                                       std::vector<block_work_output>& work_output)
    {
        return work(work_input, work_output);
    }
    void set_scheduler(std::shared_ptr<scheduler> sched) { p_scheduler = std::move(sched); }

    // Called once by the scheduler when the edge buffers exist (flowgraph::validate(), i.e. before start()
    // and outside the timed start()->wait() region): inputs[i] is the buffer of input port i, outputs[o] the
    // buffers of output port o (several when it fans out).  Not in the reference -- there a block first
    // sees its buffers inside work().  Default: nothing.
    virtual void buffers_attached(const std::vector<std::shared_ptr<buffer>>& /*inputs*/,
                                  const std::vector<std::vector<std::shared_ptr<buffer>>>& /*outputs*/)
    {
    }
};
typedef block::sptr block_sptr;
typedef std::vector<block_sptr> block_vector_t;

} // namespace gr
