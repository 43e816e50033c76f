// gnuradio/vmcircbuf.hpp -- host circular buffer with a doubly mapped window (default edge
// buffer of the `mt` scheduler, reference runtime/include/gnuradio/vmcircbuf.hpp and
// runtime/lib/vmcircbuf*.cpp).  One anonymous memory file (memfd_create) mapped twice back to
// back, so any window of <= capacity bytes is linearly addressable.  Same accounting rules as
// the reference: the write window is min(capacity - size - 1, capacity/2) items
// (vmcircbuf.cpp:79-83).
#pragma once
#include <gnuradio/buffer.hpp>

#include <sys/mman.h>
#include <unistd.h>

#include <cstring>
#include <stdexcept>

namespace gr {

enum class vmcirc_buffer_type { AUTO, SYSV_SHM, MMAP_SHM, MMAP_TMPFILE };

class vmcirc_buffer_properties : public buffer_properties
{
    vmcirc_buffer_type _type;

public:
    explicit vmcirc_buffer_properties(vmcirc_buffer_type t) : _type(t) {}
    vmcirc_buffer_type buffer_type() { return _type; }
    static std::shared_ptr<buffer_properties> make(vmcirc_buffer_type t = vmcirc_buffer_type::AUTO)
    {
        return std::make_shared<vmcirc_buffer_properties>(t);
    }
};

class vmcirc_buffer : public buffer
{
    uint8_t* _base = nullptr;
    size_t _item_size, _num_items, _buf_size; // _buf_size = bytes of ONE mapping
    size_t _read_index = 0, _write_index = 0; // bytes

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
    vmcirc_buffer(size_t num_items, size_t item_size) : _item_size(item_size)
    {
        const size_t page = (size_t)sysconf(_SC_PAGESIZE);
        const size_t unit = lcm(page, item_size);
        size_t want = std::max<size_t>(num_items, 1) * item_size;
        _buf_size = (want + unit - 1) / unit * unit;
        _num_items = _buf_size / item_size;
        int fd = memfd_create("gr_vmcirc", 0);
        if (fd < 0 || ftruncate(fd, (off_t)_buf_size) != 0)
            throw std::runtime_error("vmcirc_buffer: memfd_create/ftruncate failed");
        void* r = mmap(nullptr, 2 * _buf_size, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (r == MAP_FAILED)
            throw std::runtime_error("vmcirc_buffer: reserve failed");
        void* a = mmap(r, _buf_size, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_FIXED, fd, 0);
        void* b = mmap((uint8_t*)r + _buf_size, _buf_size, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_FIXED, fd, 0);
        close(fd);
        if (a == MAP_FAILED || b == MAP_FAILED)
            throw std::runtime_error("vmcirc_buffer: double mapping failed");
        _base = (uint8_t*)r;
        set_type("vmcirc_buffer");
    }
    ~vmcirc_buffer() override
    {
        if (_base)
            munmap(_base, 2 * _buf_size);
    }
    static buffer_sptr make(size_t num_items, size_t item_size, std::shared_ptr<buffer_properties>)
    {
        return buffer_sptr(new vmcirc_buffer(num_items, item_size));
    }

    int size() // items readable
    {
        size_t w = _write_index, r = _read_index;
        if (w < r)
            w += _buf_size;
        return (int)((w - r) / _item_size);
    }
    int capacity() { return (int)_num_items; }
    void* read_ptr() override { return _base + _read_index; }
    void* write_ptr() override { return _base + _write_index; }

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
        info.ptr = write_ptr();
        int space = capacity() - size() - 1;
        info.n_items = std::max(0, std::min(space, capacity() / 2));
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
        _write_index = (_write_index + (size_t)n * _item_size) % _buf_size;
        _total_written += n;
    }
    void copy_items(std::shared_ptr<buffer> from, int nitems) override
    {
        std::scoped_lock g(_buf_mutex);
        memcpy(write_ptr(), from->write_ptr(), (size_t)nitems * _item_size);
    }
};

} // namespace gr
