// gnuradio/buffer.hpp -- the abstract edge buffer every block reads from / writes to.
//
// API-compatible restatement of reference runtime/include/gnuradio/buffer.hpp:18-23 (buffer_info_t),
// :29-205 (buffer: read_ptr/write_ptr/read_info/write_info/post_read/post_write/copy_items + tag
// store keyed by absolute item offsets), :214-219 (buffer_properties), :221-223 (factory type).
// Additions (not in the reference, used by the drain logic of the harness scheduler, which the
// reference lacks -- SURVEY.md 7.3 "End-of-stream drain"): writer_done()/reader_done() flags.
#pragma once
#include <gnuradio/tag.hpp>

#include <atomic>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace gr {

struct buffer_info_t {
    void* ptr;
    int n_items; // items readable / writable right now, linearly addressable from ptr
    size_t item_size;
    int total_items; // items read / written so far (truncated like the reference's int)
};

class buffer
{
protected:
    std::string _name, _type;
    uint64_t _total_read = 0, _total_written = 0;
    std::mutex _buf_mutex;
    std::vector<tag_t> _tags;
    std::atomic<bool> _writer_done{ false }, _reader_done{ false };

    void set_type(const std::string& type) { _type = type; }

public:
    virtual ~buffer() {}
    virtual void* read_ptr() = 0;
    virtual void* write_ptr() = 0;
    virtual bool read_info(buffer_info_t& info) = 0;
    virtual bool write_info(buffer_info_t& info) = 0;
    virtual void post_read(int num_items) = 0;
    virtual void post_write(int num_items) = 0;
    // fan-out: copy nitems from from->write_ptr() to this->write_ptr() before either post_write
    virtual void copy_items(std::shared_ptr<buffer> from, int nitems) = 0;

    // ---- tags (offsets are absolute item counts of this edge)
    virtual std::vector<tag_t> get_tags(unsigned int num_items)
    {
        std::scoped_lock g(_buf_mutex);
        std::vector<tag_t> r;
        for (auto& t : _tags)
            if (t.offset >= _total_read && t.offset < _total_read + num_items)
                r.push_back(t);
        return r;
    }
    virtual void add_tags(unsigned int num_items, std::vector<tag_t>& tags)
    {
        std::scoped_lock g(_buf_mutex);
        for (auto& t : tags)
            if (t.offset + num_items >= _total_written && t.offset < _total_written)
                _tags.push_back(t);
    }
    // snapshot under the lock (the reference hands out an unlocked reference, buffer.hpp:73, which races
    // with add_tag / propagate_tags from the producer's thread)
    std::vector<tag_t> tags()
    {
        std::scoped_lock g(_buf_mutex);
        return _tags;
    }
    bool has_tags()
    {
        std::scoped_lock g(_buf_mutex);
        return !_tags.empty();
    }
    std::vector<tag_t> tags_in_window(uint64_t item_start, uint64_t item_end)
    {
        std::scoped_lock g(_buf_mutex);
        std::vector<tag_t> r;
        for (auto& t : _tags)
            if (t.offset >= _total_read + item_start && t.offset < _total_read + item_end)
                r.push_back(t);
        return r;
    }
    void add_tag(tag_t tag)
    {
        std::scoped_lock g(_buf_mutex);
        _tags.push_back(std::move(tag));
    }
    void add_tag(uint64_t offset, pmtf::pmt_sptr key, pmtf::pmt_sptr value, pmtf::pmt_sptr srcid = nullptr)
    {
        std::scoped_lock g(_buf_mutex);
        _tags.emplace_back(offset, key, value, srcid);
    }
    // copy the tags of the window the block just consumed from its input edge onto this (output)
    // edge; called before post_write, so the window starts at total_written() of this edge.
    // For rate-changing blocks the relative position is scaled by n_produced / n_consumed
    // ("decimation-scaled offsets"; the reference has no rate-change support).
    void propagate_tags(std::shared_ptr<buffer> in_buf, int n_consumed, int n_produced = -1)
    {
        std::vector<tag_t> src = in_buf->get_tags((unsigned)std::max(n_consumed, 0));
        uint64_t in_base = in_buf->total_read();
        std::scoped_lock g(_buf_mutex);
        for (auto& t : src) {
            tag_t c = t;
            uint64_t rel = t.offset - in_base;
            if (n_produced >= 0 && n_consumed > 0 && n_produced != n_consumed)
                rel = rel * (uint64_t)n_produced / (uint64_t)n_consumed;
            c.offset = _total_written + rel;
            _tags.push_back(c);
        }
    }
    void prune_tags(int n_consumed)
    {
        std::scoped_lock g(_buf_mutex);
        std::vector<tag_t> keep;
        for (auto& t : _tags)
            if (t.offset >= _total_read + (uint64_t)std::max(n_consumed, 0))
                keep.push_back(t);
        _tags.swap(keep);
    }

    void set_name(const std::string& name) { _name = name; }
    std::string name() { return _name; }
    std::string type() { return _type; }
    uint64_t total_written() const { return _total_written; }
    uint64_t total_read() const { return _total_read; }
    // items written but not yet read, sampled under the buffer's own lock (safe from any thread)
    uint64_t items_pending()
    {
        std::scoped_lock g(_buf_mutex);
        return _total_written - _total_read;
    }

    // end-of-stream: set by the scheduler when the block feeding this edge has finished
    void set_writer_done() { _writer_done.store(true, std::memory_order_release); }
    bool writer_done() const { return _writer_done.load(std::memory_order_acquire); }
    // ... and when the block draining this edge has finished (e.g. head returned WORK_DONE)
    void set_reader_done() { _reader_done.store(true, std::memory_order_release); }
    bool reader_done() const { return _reader_done.load(std::memory_order_acquire); }
};

typedef std::shared_ptr<buffer> buffer_sptr;

// Optional capability of an edge buffer whose far side is a device (gnuradio/devicebuffer.hpp): items can be
// handed over from / delivered into page-locked host memory OWNED BY THE BLOCK, without passing through the
// buffer's own staging ring.  Not in the reference (its cuda_buffer always stages, cudabuffer.cu:126-158);
// host blocks probe for it with dynamic_cast and fall back to read_ptr()/write_ptr() when it is absent.
class host_direct_io
{
public:
    virtual ~host_direct_io() {}
    // page-lock [p, p + bytes) for asynchronous copies; the buffer releases it when it is destroyed
    virtual bool pin_host(void* p, size_t bytes) = 0;
    // producer side (host -> device edge): this work() call's n_items output items are at `src`; the block
    // then reports n_produced as usual and does not touch write_ptr()
    virtual bool write_from_host(const void* src, int n_items) = 0;
    // consumer side (device -> host edge): deliver the stream, in order, into dst[0 .. capacity_items) and
    // offer it to the reader there (read_ptr() then points into dst); anything beyond the capacity goes
    // through the staging ring again.  dst = nullptr switches it off
    virtual bool deliver_into_host(void* dst, uint64_t capacity_items) = 0;
    static host_direct_io* from(const buffer_sptr& b) { return dynamic_cast<host_direct_io*>(b.get()); }
};

class buffer_properties
{
public:
    buffer_properties() {}
    virtual ~buffer_properties() {}
};

typedef std::function<std::shared_ptr<buffer>(size_t, size_t, std::shared_ptr<buffer_properties>)>
    buffer_factory_function;

} // namespace gr
