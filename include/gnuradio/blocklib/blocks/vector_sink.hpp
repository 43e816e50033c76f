// blocks::vector_sink<T> -- collects everything it receives (reference API:
// blocklib/blocks/include/gnuradio/blocklib/blocks/vector_sink.hpp:11-52; the reference pushes
// element by element, lib/vector_sink.cpp:29-40 -- here one insert per call).
#pragma once
#include <gnuradio/sync_block.hpp>

namespace gr {
namespace blocks {

template <class T>
class vector_sink : public sync_block
{
public:
    typedef std::shared_ptr<vector_sink> sptr;
    static sptr make(const size_t vlen = 1, const size_t reserve_items = 1024)
    {
        auto ptr = std::make_shared<vector_sink>(vlen, reserve_items);
        ptr->add_port(port<T>::make("input", port_direction_t::INPUT, std::vector<size_t>{ vlen }));
        return ptr;
    }
    vector_sink(const size_t vlen = 1, const size_t reserve_items = 1024) : sync_block("vector_sink"), d_vlen(vlen)
    {
        d_data.reserve(d_vlen * reserve_items);
    }
    work_return_code_t work(std::vector<block_work_input>& work_input, std::vector<block_work_output>&) override
    {
        const T* iptr = (const T*)work_input[0].buffer->read_ptr();
        size_t n = (size_t)work_input[0].n_items * d_vlen;
        d_data.insert(d_data.end(), iptr, iptr + n);
        auto tags = work_input[0].buffer->get_tags(work_input[0].n_items);
        d_tags.insert(d_tags.end(), tags.begin(), tags.end());
        work_input[0].n_consumed = work_input[0].n_items;
        return work_return_code_t::WORK_OK;
    }
    std::vector<T> data() { return d_data; }
    std::vector<tag_t> tags() { return d_tags; }

private:
    std::vector<T> d_data;
    std::vector<tag_t> d_tags;
    size_t d_vlen;
};
typedef vector_sink<std::uint8_t> vector_sink_b;
typedef vector_sink<std::int16_t> vector_sink_s;
typedef vector_sink<std::int32_t> vector_sink_i;
typedef vector_sink<float> vector_sink_f;
typedef vector_sink<gr_complex> vector_sink_c;

} // namespace blocks
} // namespace gr
