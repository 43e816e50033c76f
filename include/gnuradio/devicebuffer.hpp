// gnuradio/devicebuffer.hpp -- device-resident edge buffer for GPU-to-GPU edges, plus the
// host-to-device and device-to-host staging flavours.
//
// Drop-in for the reference's gr::cuda_buffer (runtime/include/gnuradio/cudabuffer.hpp:11-86,
// runtime/lib/cudabuffer.cu:17-183): same factory signature (buffer_factory_function,
// buffer.hpp:221-223), same properties pattern and the same edge->set_custom_buffer(...) macros
// (DEVICE_BUFFER_ARGS_{H2D,D2D,D2H}; gnuradio/cudabuffer.hpp maps the reference's
// CUDA_BUFFER_ARGS_* names onto them).  What is different underneath:
//   * the device ring is ONE allocation mapped twice with the CUDA VMM API (b200_ring_*), so
//     read_ptr()/write_ptr() windows are linearly addressable with no mirror copies -- the
//     reference copies every written span a second time (cudabuffer.cu:130-169);
//   * nothing synchronises the device on a D2D edge: ordering between the producer's and the
//     consumer's streams is carried by two CUDA events owned by the buffer (written / read);
//     the reference calls cudaStreamSynchronize under the buffer mutex in every post_write
//     (cudabuffer.cu:118,175);
//   * the ring is sized from the PROPERTIES (default 64 MiB), not from the scheduler's
//     2*32768-byte default (buffer_management.cpp:117), so one work() covers millions of items;
//   * _total_read/_total_written are maintained, so stream tags work on device edges
//     (the reference's cuda_buffer never updates them).
// A GPU block brackets its launches with device_stream_guard (below): wait for the producer's
// writes / the consumer's reads on ITS stream before launching, record both events after.
#pragma once
#include <gnuradio/buffer.hpp>

#include <b200dsp.h>

#include <cstring>
#include <stdexcept>

namespace gr {

enum class device_buffer_type { D2D, H2D, D2H };

// What a GPU block needs from an edge buffer to order its asynchronous launches against its
// neighbours (implemented by device_buffer below and by pinned_buffer in cudabuffer_pinned.hpp).
class stream_ordered_buffer
{
public:
    virtual ~stream_ordered_buffer() {}
    virtual void wait_readable(b200_stream_t s) = 0; // the data about to be read has been written
    virtual void wait_writable(b200_stream_t s) = 0; // the space about to be written has been read
    virtual void record_read(b200_stream_t s) = 0;
    virtual void record_write(b200_stream_t s) = 0;
    static stream_ordered_buffer* from(const buffer_sptr& b) { return dynamic_cast<stream_ordered_buffer*>(b.get()); }
};

class device_buffer_properties : public buffer_properties
{
    device_buffer_type _buffer_type;
    size_t _bytes;

public:
    static constexpr size_t default_bytes = 64u << 20;
    device_buffer_properties(device_buffer_type t, size_t bytes = default_bytes) : _buffer_type(t), _bytes(bytes) {}
    device_buffer_type buffer_type() { return _buffer_type; }
    size_t bytes() { return _bytes; }
    static std::shared_ptr<buffer_properties> make(device_buffer_type t, size_t bytes = default_bytes)
    {
        return std::make_shared<device_buffer_properties>(t, bytes);
    }
};

class device_buffer : public buffer, public stream_ordered_buffer
{
    device_buffer_type _buffer_type;
    size_t _item_size, _num_items = 0, _buf_size = 0; // bytes of one mapping
    size_t _read_index = 0, _write_index = 0;         // bytes
    b200_ring* _ring = nullptr;
    uint8_t* _dev = nullptr;
    uint8_t* _host = nullptr; // pinned staging (H2D: producer side, D2H: consumer side)
    b200_stream_t _stream = nullptr;
    b200_event_t _ev_written = nullptr, _ev_read = nullptr, _ev_copy[2] = { nullptr, nullptr };
    int _copy_slot = 0;
    bool _copy_pending[2] = { false, false };

    static void ck(int rc, const char* what)
    {
        if (rc != B200_OK)
            throw std::runtime_error(std::string("device_buffer: ") + what + ": " + b200_last_error());
    }
    static size_t lcm(size_t a, size_t b)
    {
        size_t x = a, y = b;
        while (y) {
            size_t t = x % y;
            x = y;
            y = t;
        }
        return a / x * b;
    }

public:
    typedef std::shared_ptr<device_buffer> sptr;
    device_buffer(size_t /*num_items from the scheduler: ignored*/, size_t item_size,
                  device_buffer_type type, size_t bytes)
        : _buffer_type(type), _item_size(item_size)
    {
        size_t want = std::max(bytes, 4 * item_size);
        if (type != device_buffer_type::D2D) {
            // the host side is not doubly mapped: keep the wrap point on an item boundary
            size_t gran = b200_ring_granularity();
            if (gran == 0)
                throw std::runtime_error(std::string("device_buffer: no CUDA VMM: ") + b200_last_error());
            size_t unit = lcm(gran, item_size);
            want = (want + unit - 1) / unit * unit;
        }
        ck(b200_ring_create(want, &_ring), "ring_create");
        _dev = (uint8_t*)b200_ring_base(_ring);
        _buf_size = b200_ring_size(_ring);
        _num_items = _buf_size / item_size;
        ck(b200_stream_create(&_stream), "stream_create");
        ck(b200_event_create(&_ev_written, 0), "event_create");
        ck(b200_event_create(&_ev_read, 0), "event_create");
        ck(b200_event_create(&_ev_copy[0], 0), "event_create");
        ck(b200_event_create(&_ev_copy[1], 0), "event_create");
        // make the events "complete" so the first waits fall through
        ck(b200_event_record(_ev_written, _stream), "event_record");
        ck(b200_event_record(_ev_read, _stream), "event_record");
        if (type != device_buffer_type::D2D)
            ck(b200_host_alloc((void**)&_host, _buf_size), "host_alloc");
        set_type("device_buffer_" + std::to_string((int)type));
    }
    ~device_buffer() override
    {
        if (_stream)
            b200_stream_synchronize(_stream);
        if (_host)
            b200_host_free(_host);
        for (auto e : { _ev_written, _ev_read, _ev_copy[0], _ev_copy[1] })
            if (e)
                b200_event_destroy(e);
        if (_stream)
            b200_stream_destroy(_stream);
        b200_ring_destroy(_ring);
    }

    static buffer_sptr make(size_t num_items, size_t item_size, std::shared_ptr<buffer_properties> props)
    {
        auto p = std::dynamic_pointer_cast<device_buffer_properties>(props);
        if (!p)
            throw std::runtime_error("Failed to cast buffer properties to device_buffer_properties");
        return buffer_sptr(new device_buffer(num_items, item_size, p->buffer_type(), p->bytes()));
    }
    static device_buffer* from(const buffer_sptr& b) { return dynamic_cast<device_buffer*>(b.get()); }

    device_buffer_type buffer_type() const { return _buffer_type; }
    size_t bytes() const { return _buf_size; }
    int size()
    {
        size_t w = _write_index, r = _read_index;
        if (w < r)
            w += _buf_size;
        return (int)((w - r) / _item_size);
    }
    int capacity() { return (int)_num_items; }

    // device pointer except on the host-facing side of a staging buffer
    void* read_ptr() override
    {
        return _buffer_type == device_buffer_type::D2H ? (void*)(_host + _read_index) : (void*)(_dev + _read_index);
    }
    void* write_ptr() override
    {
        return _buffer_type == device_buffer_type::H2D ? (void*)(_host + _write_index) : (void*)(_dev + _write_index);
    }

    bool read_info(buffer_info_t& info) override
    {
        std::scoped_lock g(_buf_mutex);
        info.ptr = read_ptr();
        int n = size();
        if (_buffer_type == device_buffer_type::D2H) // host side is not doubly mapped: clip at the wrap
            n = std::min<int>(n, (int)((_buf_size - _read_index) / _item_size));
        info.n_items = n;
        info.item_size = _item_size;
        info.total_items = (int)_total_read;
        return true;
    }
    bool write_info(buffer_info_t& info) override
    {
        std::scoped_lock g(_buf_mutex);
        info.ptr = write_ptr();
        int n = capacity() - size() - 1;
        // Small items (a complex sample is 8 bytes): end the offered window on a 128-byte line, so
        // that the NEXT window -- and with it the next read pointer of the consumer -- starts on
        // one.  The kernels stage 16-byte-aligned input with TMA and fall back to a slower path
        // otherwise; "capacity - size - 1" is odd and used to push every second call off it
        // (fir_filter_ccf in a flowgraph: 190 -> 230 GS/s).
        if (_item_size < 128 && 128 % _item_size == 0) {
            const int q = (int)(128 / _item_size);
            const int end = (int)((_write_index / _item_size + (size_t)n) % (size_t)q);
            if (n > end)
                n -= end;
        }
        if (_buffer_type == device_buffer_type::H2D) {
            n = std::min<int>(n, (int)((_buf_size - _write_index) / _item_size));
            n = std::min<int>(n, capacity() / 2); // keeps the in-flight H2D source span untouched
        }
        info.n_items = std::max(0, n);
        info.item_size = _item_size;
        info.total_items = (int)_total_written;
        return true;
    }
    void post_read(int n) override
    {
        std::scoped_lock g(_buf_mutex);
        _read_index = (_read_index + (size_t)n * _item_size) % _buf_size;
        _total_read += n;
    }
    void post_write(int n) override
    {
        std::scoped_lock g(_buf_mutex);
        const size_t nbytes = (size_t)n * _item_size;
        if (_buffer_type == device_buffer_type::H2D) {
            // device destination may still be read by an in-flight consumer kernel
            ck(b200_stream_wait_event(_stream, _ev_read), "wait_event");
            ck(b200_memcpy_h2d(_dev + _write_index, _host + _write_index, nbytes, _stream), "memcpy_h2d");
            ck(b200_event_record(_ev_written, _stream), "event_record");
            // allow ONE copy in flight: wait for the previous one before the producer refills
            ck(b200_event_record(_ev_copy[_copy_slot], _stream), "event_record");
            _copy_pending[_copy_slot] = true;
            _copy_slot ^= 1;
            if (_copy_pending[_copy_slot]) {
                ck(b200_event_synchronize(_ev_copy[_copy_slot]), "event_synchronize");
                _copy_pending[_copy_slot] = false;
            }
        } else if (_buffer_type == device_buffer_type::D2H) {
            // the producer block recorded _ev_written after its launches; bring the span to the
            // host and only then publish it to the (CPU) consumer
            ck(b200_stream_wait_event(_stream, _ev_written), "wait_event");
            size_t first = std::min(nbytes, _buf_size - _write_index);
            ck(b200_memcpy_d2h(_host + _write_index, _dev + _write_index, first, _stream), "memcpy_d2h");
            if (nbytes > first)
                ck(b200_memcpy_d2h(_host, _dev, nbytes - first, _stream), "memcpy_d2h");
            ck(b200_event_record(_ev_read, _stream), "event_record"); // device span free again
            ck(b200_stream_synchronize(_stream), "stream_synchronize");
        }
        _write_index = (_write_index + nbytes) % _buf_size;
        _total_written += n;
    }
    // fan-out (graph_executor.cpp:188-201): duplicate what the producer just wrote into the first
    // buffer of the port, before either side's post_write.  Host-staged producer side (H2D):
    // plain host memcpy into our own staging, our post_write then does the H2D copy.  Otherwise a
    // device-to-device copy on this buffer's stream, ordered by events.
    void copy_items(std::shared_ptr<buffer> from, int nitems) override
    {
        auto* src = from_checked(from);
        std::scoped_lock g(_buf_mutex);
        const size_t nbytes = (size_t)nitems * _item_size;
        if (_buffer_type == device_buffer_type::H2D) {
            if (src->_buffer_type != device_buffer_type::H2D)
                throw std::runtime_error("device_buffer::copy_items: mixed H2D / device fan-out");
            memcpy(_host + _write_index, src->_host + src->_write_index, nbytes);
            return;
        }
        ck(b200_stream_wait_event(_stream, src->_ev_written), "wait_event");
        ck(b200_stream_wait_event(_stream, _ev_read), "wait_event");
        ck(b200_memcpy_d2d(_dev + _write_index, src->_dev + src->_write_index, nbytes, _stream), "memcpy_d2d");
        ck(b200_event_record(_ev_written, _stream), "event_record");
    }

    // ---- stream ordering used by GPU blocks (see device_stream_guard)
    void wait_readable(b200_stream_t s) override { ck(b200_stream_wait_event(s, _ev_written), "wait_event"); }
    void wait_writable(b200_stream_t s) override { ck(b200_stream_wait_event(s, _ev_read), "wait_event"); }
    void record_read(b200_stream_t s) override { ck(b200_event_record(_ev_read, s), "event_record"); }
    void record_write(b200_stream_t s) override { ck(b200_event_record(_ev_written, s), "event_record"); }

private:
    static device_buffer* from_checked(const buffer_sptr& b)
    {
        auto* d = from(b);
        if (!d)
            throw std::runtime_error("device_buffer::copy_items: fan-out between different buffer types");
        return d;
    }
};

// RAII bracket for a GPU block's work(): before the launches make the block's stream wait for
// the data it is about to read (producer's writes) and for the space it is about to overwrite
// (consumer's reads); afterwards record both events.  Non-device buffers are ignored.
struct block_work_input;
struct block_work_output;
template <class In, class Out>
class device_stream_guard
{
    In& _in;
    Out& _out;
    b200_stream_t _s;

public:
    device_stream_guard(In& in, Out& out, b200_stream_t s) : _in(in), _out(out), _s(s)
    {
        for (auto& w : _in)
            if (auto* d = stream_ordered_buffer::from(w.buffer))
                d->wait_readable(_s);
        for (auto& w : _out)
            if (auto* d = stream_ordered_buffer::from(w.buffer))
                d->wait_writable(_s);
    }
    ~device_stream_guard()
    {
        for (auto& w : _in)
            if (auto* d = stream_ordered_buffer::from(w.buffer))
                d->record_read(_s);
        for (auto& w : _out)
            if (auto* d = stream_ordered_buffer::from(w.buffer))
                d->record_write(_s);
    }
};

} // namespace gr

#define DEVICE_BUFFER_ARGS_H2D gr::device_buffer::make, gr::device_buffer_properties::make(gr::device_buffer_type::H2D)
#define DEVICE_BUFFER_ARGS_D2H gr::device_buffer::make, gr::device_buffer_properties::make(gr::device_buffer_type::D2H)
#define DEVICE_BUFFER_ARGS_D2D gr::device_buffer::make, gr::device_buffer_properties::make(gr::device_buffer_type::D2D)
#define DEVICE_BUFFER_ARGS_SIZED(type, bytes) gr::device_buffer::make, gr::device_buffer_properties::make(gr::device_buffer_type::type, (bytes))
