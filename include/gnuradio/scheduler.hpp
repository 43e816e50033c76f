// gnuradio/scheduler.hpp -- scheduler base (reference runtime/include/gnuradio/scheduler.hpp):
// owns the default buffer factory used for edges without a custom buffer.
#pragma once
#include <gnuradio/graph.hpp>

namespace gr {

class scheduler : public std::enable_shared_from_this<scheduler>
{
protected:
    std::string _name;
    buffer_factory_function _default_buf_factory = nullptr;
    std::shared_ptr<buffer_properties> _default_buf_properties = nullptr;

public:
    explicit scheduler(const std::string& name) : _name(name) {}
    virtual ~scheduler() {}
    std::string name() { return _name; }
    void set_default_buffer_factory(buffer_factory_function f, std::shared_ptr<buffer_properties> p = nullptr)
    {
        _default_buf_factory = std::move(f);
        _default_buf_properties = std::move(p);
    }
    virtual void initialize(flat_graph_sptr fg) = 0;
    virtual void start() = 0;
    virtual void stop() = 0;
    virtual void wait() = 0;
    virtual void run()
    {
        start();
        wait();
    }
};
typedef std::shared_ptr<scheduler> scheduler_sptr;

} // namespace gr
