// gnuradio/cudabuffer_pinned.hpp -- pinned-host ("zero-copy") edge buffer: one page-locked host
// ring that CPU blocks and GPU kernels both address directly (the GPU over PCIe).  Source
// compatible with the reference's runtime/include/gnuradio/cudabuffer_pinned.hpp:12-73
// (cuda_buffer_pinned, cuda_buffer_pinned_properties, CUDA_BUFFER_PINNED_ARGS; used by
// schedulers/mt/bench/cuda/bm_copy.cpp with `-m 1`).  Like the reference it keeps a mirrored
// second half so windows are linearly addressable; unlike it, GPU work on either side is ordered
// by events instead of being assumed complete, totals are maintained (tags work), and the memory is
// released with the matching call (the reference frees cudaHostAlloc memory with cudaFree,
// runtime/lib/cudabuffer_pinned.cu:25).
#pragma once
#include <gnuradio/devicebuffer.hpp>

namespace gr {

class pinned_buffer_properties : public buffer_properties
{
    size_t _bytes;

public:
    explicit pinned_buffer_properties(size_t bytes = 0) : _bytes(bytes) {}
    size_t bytes() { return _bytes; }
    static std::shared_ptr<buffer_properties> make(size_t bytes = 0)
    {
        return std::make_shared<pinned_buffer_properties>(bytes);
    }
};

class pinned_buffer : public buffer, public stream_ordered_buffer
{
    uint8_t* _mem = nullptr; // 2 x _buf_size, second half mirrors the first
    size_t _item_size, _num_items, _buf_size;
    size_t _read_index = 0, _write_index = 0;
    b200_event_t _ev_written = nullptr, _ev_read = nullptr;
    bool _gpu_wrote = false, _gpu_read = false;

    static void ck(int rc, const char* what)
    {
        if (rc != B200_OK)
            throw std::runtime_error(std::string("pinned_buffer: ") + what + ": " + b200_last_error());
    }

public:
    pinned_buffer(size_t num_items, size_t item_size, size_t bytes) : _item_size(item_size)
    {
        _num_items = std::max<size_t>(bytes ? bytes / item_size : num_items, 4);
        _buf_size = _num_items * item_size;
        ck(b200_host_alloc((void**)&_mem, 2 * _buf_size), "host_alloc");
        ck(b200_event_create(&_ev_written, 0), "event_create");
        ck(b200_event_create(&_ev_read, 0), "event_create");
        set_type("cuda_buffer_pinned");
    }
    ~pinned_buffer() override
    {
        if (_ev_written)
            b200_event_synchronize(_ev_written);
        if (_ev_read)
            b200_event_synchronize(_ev_read);
        for (auto e : { _ev_written, _ev_read })
            if (e)
                b200_event_destroy(e);
        if (_mem)
            b200_host_free(_mem);
    }
    static buffer_sptr make(size_t num_items, size_t item_size, std::shared_ptr<buffer_properties> props)
    {
        auto p = std::dynamic_pointer_cast<pinned_buffer_properties>(props);
        return buffer_sptr(new pinned_buffer(num_items, item_size, p ? p->bytes() : 0));
    }

    int size()
    {
        size_t w = _write_index, r = _read_index;
        if (w < r)
            w += _buf_size;
        return (int)((w - r) / _item_size);
    }
    int capacity() { return (int)_num_items; }
    void* read_ptr() override { return _mem + _read_index; }
    void* write_ptr() override { return _mem + _write_index; }

    bool read_info(buffer_info_t& info) override
    {
        std::scoped_lock g(_buf_mutex);
        info.ptr = read_ptr();
        info.n_items = size();
        info.item_size = _item_size;
        info.total_items = (int)_total_read;
        return true;
    }
    bool write_info(buffer_info_t& info) override
    {
        std::scoped_lock g(_buf_mutex);
        // a GPU consumer may still be reading what the accounting already released
        if (_gpu_read)
            ck(b200_event_synchronize(_ev_read), "event_synchronize");
        info.ptr = write_ptr();
        info.n_items = std::max(0, std::min(capacity() - size() - 1, capacity() / 2));
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
        if (_gpu_wrote) // a GPU producer's kernels must have landed before the span is published
            ck(b200_event_synchronize(_ev_written), "event_synchronize");
        const size_t nbytes = (size_t)n * _item_size;
        const size_t first = std::min(nbytes, _buf_size - _write_index);
        memcpy(_mem + _buf_size + _write_index, _mem + _write_index, first); // mirror
        if (nbytes > first)
            memcpy(_mem, _mem + _buf_size, nbytes - first);
        _write_index = (_write_index + nbytes) % _buf_size;
        _total_written += n;
    }
    void copy_items(std::shared_ptr<buffer> from, int nitems) override
    {
        std::scoped_lock g(_buf_mutex);
        auto* src = dynamic_cast<pinned_buffer*>(from.get());
        if (!src)
            throw std::runtime_error("pinned_buffer::copy_items: fan-out between different buffer types");
        if (src->_gpu_wrote)
            ck(b200_event_synchronize(src->_ev_written), "event_synchronize");
        memcpy(write_ptr(), from->write_ptr(), (size_t)nitems * _item_size);
    }

    // GPU side: published data is already complete (post_write synchronised on the producer);
    // space handed to a GPU producer may still be read by a GPU consumer -> stream wait
    void wait_readable(b200_stream_t) override {}
    void wait_writable(b200_stream_t s) override
    {
        if (_gpu_read)
            ck(b200_stream_wait_event(s, _ev_read), "wait_event");
    }
    void record_read(b200_stream_t s) override
    {
        ck(b200_event_record(_ev_read, s), "event_record");
        _gpu_read = true;
    }
    void record_write(b200_stream_t s) override
    {
        ck(b200_event_record(_ev_written, s), "event_record");
        _gpu_wrote = true;
    }
};

using cuda_buffer_pinned = pinned_buffer;
using cuda_buffer_pinned_properties = pinned_buffer_properties;

} // namespace gr

#define CUDA_BUFFER_PINNED_ARGS cuda_buffer_pinned::make, cuda_buffer_pinned_properties::make()
#define PINNED_BUFFER_ARGS_SIZED(bytes) gr::pinned_buffer::make, gr::pinned_buffer_properties::make(bytes)
