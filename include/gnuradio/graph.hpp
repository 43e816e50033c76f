// gnuradio/graph.hpp -- nodes + edges (reference runtime/include/gnuradio/graph.hpp:24-63,
// runtime/lib/graph.cpp:5-76: connect() by endpoint, by port index, by port name; returns the
// edge so the caller can ->set_custom_buffer(...)).
#pragma once
#include <gnuradio/block.hpp>
#include <gnuradio/edge.hpp>

namespace gr {

class graph : public node, public std::enable_shared_from_this<graph>
{
protected:
    node_vector_t _nodes;
    edge_vector_t _edges;

public:
    typedef std::shared_ptr<graph> sptr;
    static sptr make() { return std::make_shared<graph>(); }
    graph() : node() {}
    edge_vector_t& edges() { return _edges; }

    edge_sptr connect(const node_endpoint& src, const node_endpoint& dst)
    {
        if (!src.port() || !dst.port())
            throw std::invalid_argument("connect: no such port");
        if (src.port()->direction() != port_direction_t::OUTPUT ||
            dst.port()->direction() != port_direction_t::INPUT)
            throw std::invalid_argument("connect: must go from an output port to an input port");
        for (auto& e : _edges)
            if (e->dst() == dst)
                throw std::invalid_argument("connect: input port already connected: " + dst.identifier());
        auto e = edge::make(src, dst);
        _edges.push_back(e);
        for (auto& n : { src.node(), dst.node() })
            if (std::find(_nodes.begin(), _nodes.end(), n) == _nodes.end())
                _nodes.push_back(n);
        src.port()->connect(dst.port());
        dst.port()->connect(src.port());
        return e;
    }
    edge_sptr connect(node_sptr src_node, unsigned int src_port_index, node_sptr dst_node,
                      unsigned int dst_port_index)
    {
        return connect(
            node_endpoint(src_node, src_node->get_port(src_port_index, port_type_t::STREAM, port_direction_t::OUTPUT)),
            node_endpoint(dst_node, dst_node->get_port(dst_port_index, port_type_t::STREAM, port_direction_t::INPUT)));
    }
    edge_sptr connect(node_sptr src_node, const std::string& src_port_name, node_sptr dst_node,
                      const std::string& dst_port_name)
    {
        return connect(node_endpoint(src_node, src_node->get_port(src_port_name)),
                       node_endpoint(dst_node, dst_node->get_port(dst_port_name)));
    }
    // remove one edge (graph rewriting before validate(): see gnuradio/blocklib/cuda/fusion.hpp)
    void disconnect(edge_sptr e)
    {
        _edges.erase(std::remove(_edges.begin(), _edges.end(), e), _edges.end());
        e->src().port()->disconnect(e->dst().port());
        e->dst().port()->disconnect(e->src().port());
        for (auto n : { e->src().node(), e->dst().node() }) {
            bool used = false;
            for (auto& o : _edges)
                used |= (o->src().node() == n || o->dst().node() == n);
            if (!used)
                _nodes.erase(std::remove(_nodes.begin(), _nodes.end(), n), _nodes.end());
        }
    }
    edge_vector_t out_edges(node_sptr n)
    {
        edge_vector_t r;
        for (auto& e : _edges)
            if (e->src().node() == n)
                r.push_back(e);
        return r;
    }
    edge_vector_t in_edges(node_sptr n)
    {
        edge_vector_t r;
        for (auto& e : _edges)
            if (e->dst().node() == n)
                r.push_back(e);
        return r;
    }
    node_vector_t calc_used_nodes() { return _nodes; }
    block_vector_t calc_used_blocks()
    {
        block_vector_t r;
        for (auto& n : _nodes)
            if (auto b = std::dynamic_pointer_cast<block>(n))
                r.push_back(b);
        return r;
    }
    edge_vector_t find_edge(port_sptr port)
    {
        edge_vector_t r;
        for (auto& e : _edges)
            if (e->src().port() == port || e->dst().port() == port)
                r.push_back(e);
        return r;
    }
};
typedef graph::sptr graph_sptr;
typedef graph flat_graph; // the harness has no hierarchical blocks: a graph is already flat
typedef graph_sptr flat_graph_sptr;

} // namespace gr
