// gnuradio/node.hpp -- anything that owns ports and can be connected in a graph
// (API of reference runtime/include/gnuradio/node.hpp:25-150: add_port, input/output_stream_ports,
// get_port(index, type, direction), name/alias/id).
#pragma once
#include <gnuradio/port.hpp>

#include <atomic>

namespace gr {

typedef uint32_t nodeid_t;

class node
{
protected:
    std::string d_name, d_alias;
    nodeid_t d_id = 0;
    std::vector<port_sptr> d_all_ports, d_input_ports, d_output_ports;

    static nodeid_t next_id()
    {
        static std::atomic<nodeid_t> counter{ 1 };
        return counter.fetch_add(1);
    }

public:
    void add_port(port_sptr p)
    {
        d_all_ports.push_back(p);
        if (p->direction() == port_direction_t::INPUT) {
            if (p->type() == port_type_t::STREAM)
                p->set_index((int)input_stream_ports().size());
            d_input_ports.push_back(p);
        } else if (p->direction() == port_direction_t::OUTPUT) {
            if (p->type() == port_type_t::STREAM)
                p->set_index((int)output_stream_ports().size());
            d_output_ports.push_back(p);
        }
    }

    node() : d_name("") {}
    explicit node(const std::string& name) : d_name(name), d_id(next_id()) { d_alias = name + std::to_string(d_id); }
    virtual ~node() {}
    typedef std::shared_ptr<node> sptr;

    std::vector<port_sptr>& all_ports() { return d_all_ports; }
    std::vector<port_sptr>& input_ports() { return d_input_ports; }
    std::vector<port_sptr>& output_ports() { return d_output_ports; }
    std::vector<port_sptr> input_stream_ports() { return filter(d_input_ports); }
    std::vector<port_sptr> output_stream_ports() { return filter(d_output_ports); }

    std::string& name() { return d_name; }
    std::string& alias() { return d_alias; }
    uint32_t id() { return d_id; }
    void set_alias(std::string alias) { d_alias = std::move(alias); }
    void set_id(uint32_t id) { d_id = id; }

    port_sptr get_port(const std::string& name)
    {
        for (auto& p : d_all_ports)
            if (p->name() == name)
                return p;
        return nullptr;
    }
    port_sptr get_port(unsigned int index, port_type_t type, port_direction_t direction)
    {
        for (auto& p : d_all_ports)
            if (p->type() == type && p->direction() == direction && p->index() == (int)index)
                return p;
        return nullptr;
    }

private:
    static std::vector<port_sptr> filter(const std::vector<port_sptr>& v)
    {
        std::vector<port_sptr> r;
        for (auto& p : v)
            if (p->type() == port_type_t::STREAM)
                r.push_back(p);
        return r;
    }
};
typedef node::sptr node_sptr;
typedef std::vector<node_sptr> node_vector_t;

} // namespace gr
