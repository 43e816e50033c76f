// gnuradio/blocklib/blocks/host_blocks.hpp -- the host-side harness blocks the BASELINE configs
// and the reference's tests are built from: vector_source<T>, vector_sink<T>, null_source,
// null_sink, head.  They only move bytes between std::vector / host edge buffers and carry no
// signal arithmetic (the product has no CPU signal path); class names, make() signatures and
// work() semantics follow the reference blocks so that its test sources compile unchanged:
//   vector_source  blocklib/blocks/include/gnuradio/blocklib/blocks/vector_source.hpp:12-50,
//                  lib/vector_source.cpp:39-82 (WORK_DONE when exhausted, optional repeat, tags)
//   vector_sink    .../vector_sink.hpp:11-52, lib/vector_sink.cpp:29-40
//   null_source    .../null_source.hpp:9-57   (memset of every output window)
//   null_sink      .../null_sink.hpp:41-45
//   head           .../head.hpp:10-75         (first nitems items, then WORK_DONE)
// The per-block headers of the same directory simply include this file.
#pragma once
#include <gnuradio/sync_block.hpp>

#include <cstring>

namespace gr {
namespace blocks {

namespace detail {
// every harness block is "sync_block + N input ports + M output ports of one item type"
template <class Block, class... Args>
std::shared_ptr<Block> assemble(const std::vector<port_sptr>& ports, Args&&... args)
{
    auto blk = std::make_shared<Block>(std::forward<Args>(args)...);
    for (auto& p : ports)
        blk->add_port(p);
    return blk;
}
inline port_sptr raw_port(const std::string& name, port_direction_t dir, size_t itemsize)
{
    return untyped_port::make(name, dir, itemsize);
}
template <class T>
port_sptr typed_port(const std::string& name, port_direction_t dir, size_t vlen)
{
    return port<T>::make(name, dir, std::vector<size_t>{ vlen });
}
} // namespace detail

// ------------------------------------------------------------------------------ vector_source
template <class T>
class vector_source : public sync_block
{
    std::vector<T> d_data;
    std::vector<tag_t> d_tags;
    size_t d_vlen, d_cursor = 0; // cursor counts scalars
    bool d_repeat;
    std::vector<host_direct_io*> d_direct; // every buffer of the output port takes page-locked memory directly

    void emit_tags(buffer& out, size_t first_scalar, size_t n_scalars)
    {
        const uint64_t lo = first_scalar / d_vlen, hi = (first_scalar + n_scalars) / d_vlen;
        const uint64_t base = out.total_written();
        for (auto& t : d_tags)
            if (t.offset >= lo && t.offset < hi)
                out.add_tag(base + (t.offset - lo), t.key, t.value, t.srcid);
    }

public:
    typedef std::shared_ptr<vector_source> sptr;
    static sptr make(const std::vector<T>& data, bool repeat = false, unsigned int vlen = 1,
                     const std::vector<tag_t>& tags = std::vector<tag_t>())
    {
        return detail::assemble<vector_source>({ detail::typed_port<T>("output", port_direction_t::OUTPUT, vlen) },
                                               data, repeat, vlen, tags);
    }
    vector_source(const std::vector<T>& data, bool repeat, unsigned int vlen, const std::vector<tag_t>& tags)
        : sync_block("vector_source"), d_data(data), d_tags(tags), d_vlen(vlen), d_repeat(repeat)
    {
        if (d_vlen == 0 || d_data.size() % d_vlen)
            throw std::invalid_argument("data length must be a multiple of vlen");
    }

    // Host-to-device edges (device_buffer H2D): page-lock the vector once, before start(), and let every
    // work() call hand its window over in place -- the DMA engine reads the std::vector itself, the
    // intermediate memcpy into the edge's staging ring (128 MiB for configs[0]) disappears.
    void buffers_attached(const std::vector<buffer_sptr>&, const std::vector<std::vector<buffer_sptr>>& outs) override
    {
        d_direct.clear();
        if (outs.empty() || outs[0].empty() || d_data.empty())
            return;
        std::vector<host_direct_io*> ios;
        for (auto& b : outs[0]) {
            auto* io = host_direct_io::from(b);
            if (!io || !io->write_from_host(nullptr, 0)) // probe: only host -> device edges accept
                return;
            ios.push_back(io);
        }
        if (ios[0]->pin_host(d_data.data(), d_data.size() * sizeof(T)))
            d_direct = ios;
    }

    work_return_code_t work(std::vector<block_work_input>&, std::vector<block_work_output>& wo) override
    {
        auto& o = wo[0];
        const size_t room = (size_t)o.n_items * d_vlen;
        if (!d_direct.empty() && !d_data.empty() && d_cursor < d_data.size() &&
            (d_repeat ? d_data.size() - d_cursor >= room : true)) {
            // in place: the window [d_cursor, d_cursor + run) of the page-locked vector is this call's output
            const size_t run = std::min(room, d_data.size() - d_cursor);
            if (!d_repeat)
                emit_tags(*o.buffer, d_cursor, run);
            d_direct[0]->write_from_host(&d_data[d_cursor], (int)(run / d_vlen)); // fan-out branches follow in copy_items
            d_cursor += run;
            if (d_repeat)
                d_cursor %= d_data.size();
            o.n_produced = (int)(run / d_vlen);
            return (!d_repeat && d_cursor >= d_data.size()) ? work_return_code_t::WORK_DONE : work_return_code_t::WORK_OK;
        }
        T* dst = static_cast<T*>(o.buffer->write_ptr());
        if (d_data.empty() || (!d_repeat && d_cursor >= d_data.size())) {
            o.n_produced = 0;
            return work_return_code_t::WORK_DONE;
        }
        if (d_repeat) { // wrap around the vector as often as needed
            for (size_t done = 0; done < room;) {
                const size_t run = std::min(room - done, d_data.size() - d_cursor);
                std::memcpy(dst + done, &d_data[d_cursor], run * sizeof(T));
                d_cursor = (d_cursor + run) % d_data.size();
                done += run;
            }
            o.n_produced = o.n_items;
            return work_return_code_t::WORK_OK;
        }
        const size_t run = std::min(room, d_data.size() - d_cursor);
        emit_tags(*o.buffer, d_cursor, run);
        std::memcpy(dst, &d_data[d_cursor], run * sizeof(T));
        d_cursor += run;
        o.n_produced = (int)(run / d_vlen);
        return d_cursor >= d_data.size() ? work_return_code_t::WORK_DONE : work_return_code_t::WORK_OK;
    }
};
typedef vector_source<std::uint8_t> vector_source_b;
typedef vector_source<std::int16_t> vector_source_s;
typedef vector_source<std::int32_t> vector_source_i;
typedef vector_source<float> vector_source_f;
typedef vector_source<gr_complex> vector_source_c;

// -------------------------------------------------------------------------------- vector_sink
template <class T>
class vector_sink : public sync_block
{
    std::vector<T> d_store;
    std::vector<tag_t> d_seen;
    size_t d_vlen;
    // device -> host edges (device_buffer D2H): the reserved storage is page-locked before start() and the
    // edge delivers the stream straight into it; work() then only advances d_count (no append pass)
    bool d_in_place = false;
    size_t d_count = 0;        // scalars delivered in place
    std::vector<T> d_overflow; // whatever arrives beyond the reservation (through the staging ring)

public:
    typedef std::shared_ptr<vector_sink> sptr;
    static sptr make(const size_t vlen = 1, const size_t reserve_items = 1024)
    {
        return detail::assemble<vector_sink>({ detail::typed_port<T>("input", port_direction_t::INPUT, vlen) }, vlen,
                                             reserve_items);
    }
    explicit vector_sink(size_t vlen = 1, size_t reserve_items = 1024) : sync_block("vector_sink"), d_vlen(vlen)
    {
        // reserve AND touch: a reserve alone leaves every page to be faulted in by the first append,
        // inside the flowgraph's run (16 Mi complex samples = 32768 page faults on the sink thread)
        d_store.resize(d_vlen * reserve_items);
        d_store.clear();
    }
    void buffers_attached(const std::vector<buffer_sptr>& ins, const std::vector<std::vector<buffer_sptr>>&) override
    {
        d_in_place = false;
        auto* io = ins.empty() ? nullptr : host_direct_io::from(ins[0]);
        const size_t cap = d_store.capacity();
        if (!io || cap < d_vlen)
            return;
        d_store.resize(cap / d_vlen * d_vlen);
        if (io->pin_host(d_store.data(), d_store.size() * sizeof(T)) &&
            io->deliver_into_host(d_store.data(), d_store.size() / d_vlen)) {
            d_in_place = true;
            d_count = 0;
        } else {
            d_store.clear();
        }
    }

    work_return_code_t work(std::vector<block_work_input>& wi, std::vector<block_work_output>&) override
    {
        auto& in = wi[0];
        const T* src = static_cast<const T*>(in.buffer->read_ptr());
        const size_t n = (size_t)in.n_items * d_vlen;
        if (d_in_place) {
            if (src == d_store.data() + d_count)
                d_count += n; // already where it belongs
            else
                d_overflow.insert(d_overflow.end(), src, src + n);
        } else
            d_store.insert(d_store.end(), src, src + n); // one append per call
        for (auto& t : in.buffer->get_tags((unsigned)in.n_items))
            d_seen.push_back(t);
        in.n_consumed = in.n_items;
        return work_return_code_t::WORK_OK;
    }
    std::vector<T> data()
    {
        if (!d_in_place)
            return d_store;
        std::vector<T> r(d_store.begin(), d_store.begin() + (std::ptrdiff_t)d_count);
        r.insert(r.end(), d_overflow.begin(), d_overflow.end());
        return r;
    }
    size_t size() const { return d_in_place ? d_count + d_overflow.size() : d_store.size(); }
    std::vector<tag_t> tags() { return d_seen; }
};
typedef vector_sink<std::uint8_t> vector_sink_b;
typedef vector_sink<std::int16_t> vector_sink_s;
typedef vector_sink<std::int32_t> vector_sink_i;
typedef vector_sink<float> vector_sink_f;
typedef vector_sink<gr_complex> vector_sink_c;

// ------------------------------------------------------------------- null_source / null_sink
class null_source : public sync_block
{
    size_t d_bytes_per_item;

public:
    typedef std::shared_ptr<null_source> sptr;
    static sptr make(size_t itemsize, size_t nports = 1)
    {
        std::vector<port_sptr> ports;
        for (size_t i = 0; i < nports; i++)
            ports.push_back(detail::raw_port("out" + std::to_string(i), port_direction_t::OUTPUT, itemsize));
        return detail::assemble<null_source>(ports, itemsize, nports);
    }
    null_source(size_t itemsize, size_t /*nports*/) : sync_block("null_source"), d_bytes_per_item(itemsize) {}
    work_return_code_t work(std::vector<block_work_input>&, std::vector<block_work_output>& wo) override
    {
        for (auto& o : wo) { // zeros into every output window (host buffers)
            std::memset(o.buffer->write_ptr(), 0, (size_t)o.n_items * d_bytes_per_item);
            o.n_produced = o.n_items;
        }
        return work_return_code_t::WORK_OK;
    }
};

class null_sink : public sync_block
{
    uint64_t d_count = 0;

public:
    typedef std::shared_ptr<null_sink> sptr;
    static sptr make(size_t itemsize, size_t nports = 1)
    {
        std::vector<port_sptr> ports;
        for (size_t i = 0; i < nports; i++)
            ports.push_back(detail::raw_port("in" + std::to_string(i), port_direction_t::INPUT, itemsize));
        return detail::assemble<null_sink>(ports, itemsize, nports);
    }
    null_sink(size_t /*itemsize*/, size_t /*nports*/) : sync_block("null_sink") {}
    // never dereferences the data, so it is equally at home on host and device edges
    work_return_code_t work(std::vector<block_work_input>& wi, std::vector<block_work_output>&) override
    {
        for (auto& in : wi) {
            d_count += (uint64_t)in.n_items;
            in.n_consumed = in.n_items;
        }
        return work_return_code_t::WORK_OK;
    }
    uint64_t n_items() const { return d_count; }
};

// --------------------------------------------------------------------------------------- head
class head : public sync_block
{
    size_t d_bytes_per_item, d_limit, d_passed = 0;

public:
    typedef std::shared_ptr<head> sptr;
    static sptr make(size_t itemsize, size_t nitems)
    {
        return detail::assemble<head>({ detail::raw_port("input", port_direction_t::INPUT, itemsize),
                                        detail::raw_port("output", port_direction_t::OUTPUT, itemsize) },
                                      itemsize, nitems);
    }
    head(size_t itemsize, size_t nitems) : sync_block("head"), d_bytes_per_item(itemsize), d_limit(nitems) {}
    work_return_code_t work(std::vector<block_work_input>& wi, std::vector<block_work_output>& wo) override
    {
        const size_t left = d_limit > d_passed ? d_limit - d_passed : 0;
        const size_t n = std::min<size_t>(left, (size_t)wo[0].n_items);
        if (n)
            std::memcpy(wo[0].buffer->write_ptr(), wi[0].buffer->read_ptr(), n * d_bytes_per_item);
        d_passed += n;
        wo[0].n_produced = (int)n;
        return d_passed >= d_limit ? work_return_code_t::WORK_DONE : work_return_code_t::WORK_OK;
    }
};

} // namespace blocks
} // namespace gr
