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
//     consumer's streams is carried by CUDA events owned by the buffer, each tagged with the item
//     position it covers, so a waiter waits for exactly the lap of the ring it is about to touch;
//     the reference calls cudaStreamSynchronize under the buffer mutex in every post_write
//     (cudabuffer.cu:118,175);
//   * the ring is sized from the PROPERTIES (default 64 MiB), not from the scheduler's
//     2*32768-byte default (buffer_management.cpp:117), so one work() covers millions of items;
//   * _total_read/_total_written are maintained, so stream tags work on device edges
//     (the reference's cuda_buffer never updates them);
//   * the staging flavours are asynchronous and multi-buffered: H2D keeps three copies in flight, D2H
//     publishes a span to the CPU reader from its completion event and never blocks the producer; host
//     blocks that own page-locked memory (vector_source / vector_sink) hand it over through
//     gr::host_direct_io, so configs[0] moves every byte exactly once in each direction.
// A GPU block brackets its launches with device_stream_guard (below): wait for the producer's
// writes / the consumer's reads on ITS stream before launching, record both events after.
#pragma once
#include <gnuradio/buffer.hpp>

#include <b200dsp.h>

#include <cstring>
#include <deque>
#include <stdexcept>
#include <vector>

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
    int _device; // -1: the device current when the buffer is created (flowgraph::validate())

public:
    static constexpr size_t default_bytes = 64u << 20;
    device_buffer_properties(device_buffer_type t, size_t bytes = default_bytes, int device = -1)
        : _buffer_type(t), _bytes(bytes), _device(device)
    {
    }
    device_buffer_type buffer_type() { return _buffer_type; }
    size_t bytes() { return _bytes; }
    int device() { return _device; }
    static std::shared_ptr<buffer_properties> make(device_buffer_type t, size_t bytes = default_bytes, int device = -1)
    {
        return std::make_shared<device_buffer_properties>(t, bytes, device);
    }
};

// the calling thread's current device for the lifetime of the object (buffers of a multi-GPU flowgraph are all
// created by the thread that calls validate())
class device_scope
{
    int _prev = -1;

public:
    explicit device_scope(int device)
    {
        if (device >= 0 && b200_get_device(&_prev) == B200_OK && _prev != device) {
            if (b200_set_device(device) != B200_OK)
                throw std::runtime_error(std::string("device_scope: ") + b200_last_error());
        } else
            _prev = -1;
    }
    ~device_scope()
    {
        if (_prev >= 0)
            b200_set_device(_prev);
    }
};

class device_buffer : public buffer, public stream_ordered_buffer, public host_direct_io
{
    device_buffer_type _buffer_type;
    size_t _item_size, _num_items = 0, _buf_size = 0; // bytes of one mapping
    size_t _read_index = 0, _write_index = 0;         // bytes; _write_index = the PRODUCER's position
    b200_ring* _ring = nullptr;
    uint8_t* _dev = nullptr;
    uint8_t* _host = nullptr; // pinned staging (H2D: producer side, D2H: consumer side)
    b200_stream_t _stream = nullptr;

    // ---- ordering across streams, by POSITION.  One event per direction ("all writes / all reads enqueued so
    // far") orders more than the ring needs: with several chunks in flight the H2D copy of chunk k+1 would wait
    // for the kernel that read chunk k, which waits for the copy of chunk k -- copies and kernels take turns and
    // the PCIe link idles between them (configs[0]: 2.6 GS/s).  What a writer of [W, W+n) really needs is that the
    // reads of the items that occupied that space ONE LAP AGO are done, i.e. reads up to item W+n-capacity; a
    // reader of [R, R+n) needs the writes up to R+n.  Every recorded event is therefore kept with the cumulative
    // item count it covers, and a waiter picks the OLDEST event that covers its position (usually long complete).
    struct mark {
        b200_event_t ev = nullptr;
        uint64_t pos = 0;
    };
    static constexpr int NMARK = 32;
    mark _wmarks[NMARK], _rmarks[NMARK]; // circular, newest at (_wn - 1) % NMARK
    uint64_t _wn = 0, _rn = 0;
    b200_event_t _w_open = nullptr, _r_open = nullptr; // recorded by the block's guard, position known at post_*()
    uint64_t _offer_w = 0, _offer_r = 0;               // windows handed out by the last write_info / read_info

    void push_mark(mark* ring, uint64_t& n, b200_event_t ev, uint64_t pos)
    {
        mark& m = ring[n % NMARK];
        if (m.ev)
            _free_events.push_back(m.ev); // oldest one falls out; waiters then use the next older-than-needed one
        m.ev = ev;
        m.pos = pos;
        n++;
    }
    // oldest event that covers `pos` (nullptr: nothing recorded covers it -- nobody touched those items on a GPU)
    b200_event_t covering(const mark* ring, uint64_t n, uint64_t pos) const
    {
        const uint64_t first = n > (uint64_t)NMARK ? n - NMARK : 0;
        for (uint64_t i = first; i < n; i++)
            if (ring[i % NMARK].pos >= pos)
                return ring[i % NMARK].ev;
        return nullptr;
    }

    // ---- H2D: up to H2D_SLOTS - 1 copies in flight; an offered window is at most capacity / H2D_SLOTS
    // items, so the staging spans of the copies still in flight are never the ones being refilled
    static constexpr int H2D_SLOTS = 4;
    b200_event_t _ev_copy[H2D_SLOTS] = { nullptr, nullptr, nullptr, nullptr };
    bool _copy_pending[H2D_SLOTS] = { false, false, false, false };
    int _copy_slot = 0;
    const uint8_t* _ext_src = nullptr; // H2D: this work() call's items sit in caller-owned pinned memory

    // ---- D2H: copies are enqueued by the producer's post_write and PUBLISHED to the (CPU) reader when
    // their completion event has fired -- the reader's thread waits for the oldest copy if it has
    // nothing else to read, the producer's thread never waits (the reference, and round 1 of this file,
    // synchronised the stream under the buffer mutex in every post_write: cudabuffer.cu:148-158)
    struct span {
        b200_event_t ev;
        uint8_t* host;
        uint64_t n_items;
    };
    std::deque<span> _pending; // copies in flight, oldest first          (both threads, under _buf_mutex)
    std::deque<span> _ready;   // landed and published, not yet consumed  (both threads, under _buf_mutex)
    std::vector<b200_event_t> _free_events;
    uint8_t* _land = nullptr;  // optional landing zone of the consumer (deliver_into_host)
    uint64_t _land_items = 0, _land_used = 0;
    std::vector<void*> _pinned; // caller memory page-locked through pin_host(), released with the buffer
    int _device = -1;

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
    b200_event_t get_event()
    {
        if (!_free_events.empty()) {
            b200_event_t e = _free_events.back();
            _free_events.pop_back();
            return e;
        }
        b200_event_t e = nullptr;
        ck(b200_event_create(&e, 0), "event_create");
        return e;
    }
    // move every landed copy from _pending to _ready (caller holds _buf_mutex)
    void publish_landed()
    {
        while (!_pending.empty()) {
            int q = b200_event_query(_pending.front().ev);
            if (q < 0)
                ck(q, "event_query");
            if (q != 0)
                break;
            span sp = _pending.front();
            _pending.pop_front();
            _free_events.push_back(sp.ev);
            sp.ev = nullptr;
            if (!_ready.empty() && _ready.back().host + _ready.back().n_items * _item_size == sp.host)
                _ready.back().n_items += sp.n_items; // contiguous in host memory: one window for the reader
            else
                _ready.push_back(sp);
        }
    }
    void enqueue_d2h(uint8_t* host_dst, const uint8_t* dev_src, uint64_t n_items)
    {
        ck(b200_memcpy_d2h(host_dst, dev_src, n_items * _item_size, _stream), "memcpy_d2h");
        span sp{ get_event(), host_dst, n_items };
        ck(b200_event_record(sp.ev, _stream), "event_record");
        _pending.push_back(sp);
    }

public:
    typedef std::shared_ptr<device_buffer> sptr;
    device_buffer(size_t /*num_items from the scheduler: ignored*/, size_t item_size,
                  device_buffer_type type, size_t bytes, int device = -1)
        : _buffer_type(type), _item_size(item_size)
    {
        device_scope on(device); // ring, stream, events and staging all belong to this device
        if (b200_get_device(&_device) != B200_OK)
            _device = -1;
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
        for (auto& e : _ev_copy)
            ck(b200_event_create(&e, 0), "event_create");
        if (type != device_buffer_type::D2D)
            ck(b200_host_alloc((void**)&_host, _buf_size), "host_alloc");
        set_type("device_buffer_" + std::to_string((int)type));
    }
    ~device_buffer() override
    {
        device_scope on(_device);
        if (_stream)
            b200_stream_synchronize(_stream);
        for (void* p : _pinned)
            b200_host_unregister(p);
        if (_host)
            b200_host_free(_host);
        for (auto& sp : _pending)
            if (sp.ev)
                b200_event_destroy(sp.ev);
        for (auto& m : _wmarks)
            if (m.ev)
                b200_event_destroy(m.ev);
        for (auto& m : _rmarks)
            if (m.ev)
                b200_event_destroy(m.ev);
        for (auto e : { _w_open, _r_open })
            if (e)
                b200_event_destroy(e);
        for (auto e : _free_events)
            b200_event_destroy(e);
        for (auto e : { _ev_copy[0], _ev_copy[1], _ev_copy[2], _ev_copy[3] })
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
        return buffer_sptr(new device_buffer(num_items, item_size, p->buffer_type(), p->bytes(), p->device()));
    }
    static device_buffer* from(const buffer_sptr& b) { return dynamic_cast<device_buffer*>(b.get()); }

    device_buffer_type buffer_type() const { return _buffer_type; }
    size_t bytes() const { return _buf_size; }
    int device() const { return _device; }
    // lets another GPU of this process read / write the ring directly (peer loads over NVLink, peer copies)
    void enable_peer(int peer_device) { ck(b200_ring_enable_peer(_ring, peer_device), "ring_enable_peer"); }
    void* device_base() const { return _dev; }
    // items between the reader's and the PRODUCER's position (what occupies the ring)
    int size() { return (int)(_total_written - _total_read); }
    int capacity() { return (int)_num_items; }

    // device pointer except on the host-facing side of a staging buffer
    void* read_ptr() override
    {
        if (_buffer_type != device_buffer_type::D2H)
            return (void*)(_dev + _read_index);
        std::scoped_lock g(_buf_mutex);
        return _ready.empty() ? (void*)(_host + _read_index) : (void*)_ready.front().host;
    }
    void* write_ptr() override
    {
        return _buffer_type == device_buffer_type::H2D ? (void*)(_host + _write_index) : (void*)(_dev + _write_index);
    }

    bool read_info(buffer_info_t& info) override
    {
        std::unique_lock<std::mutex> g(_buf_mutex);
        info.item_size = _item_size;
        info.total_items = (int)_total_read;
        if (_buffer_type != device_buffer_type::D2H) {
            info.ptr = (void*)(_dev + _read_index);
            info.n_items = size();
            _offer_r = (uint64_t)info.n_items;
            return true;
        }
        publish_landed();
        if (_ready.empty() && !_pending.empty()) {
            // nothing to hand out yet, but a copy is on its way: wait for it HERE, on the reader's thread
            // and outside the lock, so the producer keeps launching kernels and enqueuing copies meanwhile
            b200_event_t ev = _pending.front().ev;
            g.unlock();
            ck(b200_event_synchronize(ev), "event_synchronize");
            g.lock();
            publish_landed();
        }
        if (_ready.empty()) {
            info.ptr = (void*)(_host + _read_index);
            info.n_items = 0;
        } else {
            info.ptr = (void*)_ready.front().host;
            info.n_items = (int)std::min<uint64_t>(_ready.front().n_items, 0x7fffffff);
        }
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
            n = std::min<int>(n, capacity() / H2D_SLOTS); // keeps the in-flight H2D source spans untouched
        }
        info.n_items = std::max(0, n);
        _offer_w = (uint64_t)info.n_items;
        info.item_size = _item_size;
        info.total_items = (int)_total_written;
        return true;
    }
    void post_read(int n) override
    {
        std::scoped_lock g(_buf_mutex);
        if (_r_open) { // the reading block's event now has a position: reads up to here are done when it fires
            if (n > 0)
                push_mark(_rmarks, _rn, _r_open, _total_read + (uint64_t)n);
            else
                _free_events.push_back(_r_open);
            _r_open = nullptr;
        }
        if (_buffer_type == device_buffer_type::D2H) {
            uint64_t left = (uint64_t)n;
            while (left && !_ready.empty()) {
                span& f = _ready.front();
                const uint64_t take = std::min(left, f.n_items);
                f.host += take * _item_size;
                f.n_items -= take;
                left -= take;
                if (f.n_items == 0)
                    _ready.pop_front();
            }
        }
        _read_index = (_read_index + (size_t)n * _item_size) % _buf_size;
        _total_read += n;
    }
    void post_write(int n) override
    {
        std::scoped_lock g(_buf_mutex);
        const size_t nbytes = (size_t)n * _item_size;
        if (_buffer_type == device_buffer_type::H2D) {
            // the device destination may still be read by a consumer kernel -- the one that read this space a lap ago
            if (_total_written + (uint64_t)n > _num_items)
                if (b200_event_t e = covering(_rmarks, _rn, _total_written + (uint64_t)n - _num_items))
                    ck(b200_stream_wait_event(_stream, e), "wait_event");
            b200_event_t wev = get_event();
            if (_ext_src) {
                // the block handed over caller-owned page-locked memory (write_from_host): copy from there,
                // nothing of ours is reused, so there is nothing to throttle
                ck(b200_memcpy_h2d(_dev + _write_index, _ext_src, nbytes, _stream), "memcpy_h2d");
                ck(b200_event_record(wev, _stream), "event_record");
                _ext_src = nullptr;
            } else {
                ck(b200_memcpy_h2d(_dev + _write_index, _host + _write_index, nbytes, _stream), "memcpy_h2d");
                ck(b200_event_record(wev, _stream), "event_record");
                // H2D_SLOTS - 1 copies stay in flight: before the producer refills, wait for the copy that
                // was enqueued H2D_SLOTS - 1 calls ago (its staging span is the next one to be reused)
                ck(b200_event_record(_ev_copy[_copy_slot], _stream), "event_record");
                _copy_pending[_copy_slot] = true;
                _copy_slot = (_copy_slot + 1) % H2D_SLOTS;
                if (_copy_pending[_copy_slot]) {
                    ck(b200_event_synchronize(_ev_copy[_copy_slot]), "event_synchronize");
                    _copy_pending[_copy_slot] = false;
                }
            }
            push_mark(_wmarks, _wn, wev, _total_written + (uint64_t)n);
        } else if (_buffer_type == device_buffer_type::D2H) {
            // the producer block recorded its event after its launches: enqueue the copy behind it and
            // return; the span becomes visible to the reader when its completion event has fired
            if (_w_open) {
                ck(b200_stream_wait_event(_stream, _w_open), "wait_event");
                _free_events.push_back(_w_open);
                _w_open = nullptr;
            }
            uint64_t left = (uint64_t)n;
            size_t src = _write_index;
            if (_land && _land_used < _land_items) { // straight into the consumer's own storage
                const uint64_t take = std::min(left, _land_items - _land_used);
                enqueue_d2h(_land + _land_used * _item_size, _dev + src, take); // device ring is doubly mapped
                _land_used += take;
                left -= take;
                src = (src + take * _item_size) % _buf_size;
            }
            while (left) { // staging ring: split at the wrap of the (singly mapped) host side
                const uint64_t take = std::min<uint64_t>(left, (_buf_size - src) / _item_size);
                enqueue_d2h(_host + src, _dev + src, take);
                left -= take;
                src = (src + take * _item_size) % _buf_size;
            }
            // the device span is free again once these copies are done: that IS the read of this edge's device side
            b200_event_t rev = get_event();
            ck(b200_event_record(rev, _stream), "event_record");
            push_mark(_rmarks, _rn, rev, _total_written + (uint64_t)n);
        } else if (_w_open) { // D2D: the producing block's event now has a position
            if (n > 0)
                push_mark(_wmarks, _wn, _w_open, _total_written + (uint64_t)n);
            else
                _free_events.push_back(_w_open);
            _w_open = nullptr;
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
            if (src->_ext_src) { // the producer's pinned memory feeds every branch directly
                _ext_src = src->_ext_src;
                return;
            }
            memcpy(_host + _write_index, src->_host + src->_write_index, nbytes);
            return;
        }
        // source: what the producing block just launched (its open event); destination: readers of the space a lap ago
        {
            std::scoped_lock gs(src->_buf_mutex);
            if (src->_w_open)
                ck(b200_stream_wait_event(_stream, src->_w_open), "wait_event");
        }
        if (_total_written + (uint64_t)nitems > _num_items)
            if (b200_event_t e = covering(_rmarks, _rn, _total_written + (uint64_t)nitems - _num_items))
                ck(b200_stream_wait_event(_stream, e), "wait_event");
        ck(b200_memcpy_d2d(_dev + _write_index, src->_dev + src->_write_index, nbytes, _stream), "memcpy_d2d");
        if (_w_open)
            _free_events.push_back(_w_open);
        _w_open = get_event();
        ck(b200_event_record(_w_open, _stream), "event_record"); // gets its position in our post_write
    }

    // ---- host_direct_io (gnuradio/buffer.hpp): zero-copy hand-over with host blocks
    bool pin_host(void* p, size_t bytes) override
    {
        if (!p || !bytes || b200_host_register(p, bytes) != B200_OK)
            return false;
        std::scoped_lock g(_buf_mutex);
        _pinned.push_back(p);
        return true;
    }
    bool write_from_host(const void* src, int /*n_items*/) override
    {
        if (_buffer_type != device_buffer_type::H2D)
            return false;
        std::scoped_lock g(_buf_mutex);
        _ext_src = static_cast<const uint8_t*>(src);
        return true;
    }
    bool deliver_into_host(void* dst, uint64_t capacity_items) override
    {
        if (_buffer_type != device_buffer_type::D2H)
            return false;
        std::scoped_lock g(_buf_mutex);
        _land = static_cast<uint8_t*>(dst);
        _land_items = dst ? capacity_items : 0;
        _land_used = 0;
        return true;
    }

    // ---- stream ordering used by GPU blocks (see device_stream_guard)
    // the block is about to read the window read_info() offered: writes up to its end must have landed
    void wait_readable(b200_stream_t s) override
    {
        std::scoped_lock g(_buf_mutex);
        if (b200_event_t e = covering(_wmarks, _wn, _total_read + _offer_r))
            ck(b200_stream_wait_event(s, e), "wait_event");
    }
    // ... and to write (at most) the window write_info() offered: reads of what sat there a lap ago must be done
    void wait_writable(b200_stream_t s) override
    {
        std::scoped_lock g(_buf_mutex);
        const uint64_t end = _total_written + _offer_w;
        if (end > _num_items)
            if (b200_event_t e = covering(_rmarks, _rn, end - _num_items))
                ck(b200_stream_wait_event(s, e), "wait_event");
    }
    void record_read(b200_stream_t s) override
    {
        std::scoped_lock g(_buf_mutex);
        if (!_r_open)
            _r_open = get_event();
        ck(b200_event_record(_r_open, s), "event_record"); // position assigned in post_read()
    }
    void record_write(b200_stream_t s) override
    {
        std::scoped_lock g(_buf_mutex);
        if (!_w_open)
            _w_open = get_event();
        ck(b200_event_record(_w_open, s), "event_record"); // position assigned in post_write()
    }

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
        // the block's stream decides which GPU this thread talks to: one process can drive several devices
        // (each block was built with its device current; its scheduler thread follows it here)
        if (b200_stream_activate(_s) != B200_OK)
            throw std::runtime_error(std::string("device_stream_guard: ") + b200_last_error());
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
#define DEVICE_BUFFER_ARGS_ON(type, bytes, device) gr::device_buffer::make, gr::device_buffer_properties::make(gr::device_buffer_type::type, (bytes), (device))
