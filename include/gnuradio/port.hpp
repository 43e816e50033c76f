// gnuradio/port.hpp -- typed / untyped stream ports.
// Same construction API as reference runtime/include/gnuradio/port.hpp:27-150 (port_base),
// :165-190 (port<T>::make(name, direction, dims)), :200-218 (untyped_port::make(name, direction,
// itemsize)); item size = sizeof(T) * prod(dims) (port.hpp:57-65).
#pragma once
#include <gnuradio/scheduler_message.hpp>
#include <gnuradio/types.hpp>

#include <algorithm>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace gr {

enum class port_type_t { STREAM, MESSAGE };
enum class port_direction_t { INPUT, OUTPUT, BIDIRECTONAL };

class port_base
{
public:
    typedef std::shared_ptr<port_base> sptr;
    port_base(const std::string& name, port_direction_t direction, size_t datasize,
              const std::vector<size_t>& dims, port_type_t type = port_type_t::STREAM)
        : _name(name), _direction(direction), _port_type(type), _dims(dims), _datasize(datasize),
          _itemsize(datasize)
    {
        for (auto d : _dims)
            _itemsize *= d;
    }
    virtual ~port_base() = default;

    std::string name() { return _name; }
    std::string alias() { return _alias; }
    void set_alias(const std::string& a) { _alias = a; }
    void set_index(int v) { _index = v; }
    int index() { return _index; }
    port_type_t type() { return _port_type; }
    port_direction_t direction() { return _direction; }
    size_t data_size() { return _datasize; }
    size_t itemsize() { return _itemsize; }
    std::vector<size_t> dims() { return _dims; }

    // Non-owning links.  The reference keeps shared_ptrs here (port.hpp:92-133), which closes two ownership
    // cycles -- port <-> connected port, and thread_wrapper -> block -> port -> thread_wrapper -- so a
    // flowgraph is never freed; with 64 MiB device rings + pinned staging per edge that leak matters here.
    // The scheduler owns the thread wrappers and the blocks own their ports for as long as a notification
    // can be in flight, so weak references are enough.
    void set_parent_intf(neighbor_interface_sptr intf) { _parent_intf = intf; }
    void notify_connected_ports(scheduler_message_sptr msg)
    {
        for (auto& w : _connected_ports)
            if (auto p = w.lock())
                p->push_message(msg);
    }
    virtual void push_message(scheduler_message_sptr msg)
    {
        auto intf = _parent_intf.lock();
        if (!intf)
            throw std::runtime_error("port has no parent interface");
        intf->push_message(msg);
    }
    void connect(sptr other)
    {
        for (auto& w : _connected_ports)
            if (w.lock() == other)
                return;
        _connected_ports.push_back(other);
    }

    void disconnect(sptr other)
    {
        _connected_ports.erase(std::remove_if(_connected_ports.begin(), _connected_ports.end(),
                                              [&](const std::weak_ptr<port_base>& w) {
                                                  auto p = w.lock();
                                                  return !p || p == other;
                                              }),
                               _connected_ports.end());
    }

protected:
    std::string _name, _alias;
    port_direction_t _direction;
    port_type_t _port_type;
    int _index = -1;
    std::vector<size_t> _dims;
    size_t _datasize, _itemsize;
    std::vector<std::weak_ptr<port_base>> _connected_ports;
    std::weak_ptr<neighbor_interface> _parent_intf;
};
typedef port_base::sptr port_sptr;
typedef std::vector<port_sptr> port_vector_t;

template <class T>
class port : public port_base
{
public:
    static std::shared_ptr<port<T>> make(const std::string& name, port_direction_t direction,
                                         const std::vector<size_t>& dims = std::vector<size_t>(),
                                         int /*multiplicity*/ = 1)
    {
        return std::make_shared<port<T>>(name, direction, dims);
    }
    port(const std::string& name, port_direction_t direction, const std::vector<size_t>& dims)
        : port_base(name, direction, sizeof(T), dims)
    {
    }
};

class untyped_port : public port_base
{
public:
    static std::shared_ptr<untyped_port> make(const std::string& name, port_direction_t direction,
                                              size_t itemsize, int /*multiplicity*/ = 1)
    {
        return std::make_shared<untyped_port>(name, direction, itemsize);
    }
    untyped_port(const std::string& name, port_direction_t direction, size_t itemsize)
        : port_base(name, direction, itemsize, std::vector<size_t>())
    {
    }
};

} // namespace gr
