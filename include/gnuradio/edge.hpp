// gnuradio/edge.hpp -- a connection between two ports, optionally carrying a custom buffer type.
// API of reference runtime/include/gnuradio/edge.hpp:33-107; the hook that matters for the hot
// path is set_custom_buffer(factory, properties) (:97-102), consumed by the scheduler's buffer
// manager (schedulers/mt/lib/buffer_management.cpp:78-82).
#pragma once
#include <gnuradio/buffer.hpp>
#include <gnuradio/node.hpp>

namespace gr {

class node_endpoint
{
    node_sptr d_node;
    port_sptr d_port;

public:
    node_endpoint() {}
    node_endpoint(node_sptr n, port_sptr p) : d_node(std::move(n)), d_port(std::move(p)) {}
    node_sptr node() const { return d_node; }
    port_sptr port() const { return d_port; }
    std::string identifier() const { return d_node->alias() + ":" + d_port->name(); }
};
inline bool operator==(const node_endpoint& a, const node_endpoint& b)
{
    return a.node() == b.node() && a.port() == b.port();
}

class edge
{
protected:
    node_endpoint _src, _dst;
    buffer_factory_function _buffer_factory = nullptr;
    std::shared_ptr<buffer_properties> _buffer_properties = nullptr;

public:
    typedef std::shared_ptr<edge> sptr;
    static sptr make(const node_endpoint& src, const node_endpoint& dst) { return std::make_shared<edge>(src, dst); }
    static sptr make(node_sptr sb, port_sptr sp, node_sptr db, port_sptr dp)
    {
        return std::make_shared<edge>(node_endpoint(sb, sp), node_endpoint(db, dp));
    }
    edge(const node_endpoint& src, const node_endpoint& dst) : _src(src), _dst(dst) {}
    virtual ~edge() {}
    node_endpoint src() const { return _src; }
    node_endpoint dst() const { return _dst; }
    std::string identifier() const { return _src.identifier() + "->" + _dst.identifier(); }
    size_t itemsize() const { return _src.port()->itemsize(); } // the SOURCE port decides (edge.cpp:39)

    void set_custom_buffer(buffer_factory_function factory,
                           std::shared_ptr<buffer_properties> props = nullptr)
    {
        _buffer_factory = std::move(factory);
        _buffer_properties = std::move(props);
    }
    bool has_custom_buffer() { return (bool)_buffer_factory; }
    buffer_factory_function buffer_factory() { return _buffer_factory; }
    std::shared_ptr<buffer_properties> buf_properties() { return _buffer_properties; }
};
typedef edge::sptr edge_sptr;
typedef std::vector<edge_sptr> edge_vector_t;

} // namespace gr
