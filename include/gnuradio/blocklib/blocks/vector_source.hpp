// blocks::vector_source<T> -- emits a std::vector once (or repeatedly), WORK_DONE when exhausted.
// Harness block with the reference API (blocklib/blocks/include/gnuradio/blocklib/blocks/
// vector_source.hpp:12-50, lib/vector_source.cpp:39-82); the copy is one memcpy per call.
#pragma once
#include <gnuradio/sync_block.hpp>

#include <cstring>

namespace gr {
namespace blocks {

template <class T>
class vector_source : public sync_block
{
public:
    typedef std::shared_ptr<vector_source> sptr;
    static sptr make(const std::vector<T>& data, bool repeat = false, unsigned int vlen = 1,
                     const std::vector<tag_t>& tags = std::vector<tag_t>())
    {
        auto ptr = std::make_shared<vector_source>(data, repeat, vlen, tags);
        ptr->add_port(port<T>::make("output", port_direction_t::OUTPUT, std::vector<size_t>{ vlen }));
        return ptr;
    }
    vector_source(const std::vector<T>& data, bool repeat, unsigned int vlen, const std::vector<tag_t>& tags)
        : sync_block("vector_source"), d_data(data), d_repeat(repeat), d_offset(0), d_vlen(vlen), d_tags(tags)
    {
        if (data.size() % vlen != 0)
            throw std::invalid_argument("data length must be a multiple of vlen");
    }

    work_return_code_t work(std::vector<block_work_input>&, std::vector<block_work_output>& work_output) override
    {
        T* optr = (T*)work_output[0].buffer->write_ptr();
        size_t space = (size_t)work_output[0].n_items * d_vlen; // scalars
        if (d_repeat) {
            if (d_data.empty()) {
                work_output[0].n_produced = 0;
                return work_return_code_t::WORK_DONE;
            }
            for (size_t i = 0; i < space;) {
                size_t n = std::min(space - i, d_data.size() - d_offset);
                memcpy(optr + i, d_data.data() + d_offset, n * sizeof(T));
                d_offset = (d_offset + n) % d_data.size();
                i += n;
            }
            work_output[0].n_produced = work_output[0].n_items;
            return work_return_code_t::WORK_OK;
        }
        if (d_offset >= d_data.size()) {
            work_output[0].n_produced = 0;
            return work_return_code_t::WORK_DONE;
        }
        size_t n = std::min(d_data.size() - d_offset, space);
        uint64_t first_item = work_output[0].buffer->total_written();
        for (auto& t : d_tags)
            if (t.offset >= d_offset / d_vlen && t.offset < (d_offset + n) / d_vlen)
                work_output[0].buffer->add_tag(first_item + (t.offset - d_offset / d_vlen), t.key, t.value, t.srcid);
        memcpy(optr, d_data.data() + d_offset, n * sizeof(T));
        d_offset += n;
        work_output[0].n_produced = (int)(n / d_vlen);
        return d_offset >= d_data.size() ? work_return_code_t::WORK_DONE : work_return_code_t::WORK_OK;
    }

private:
    std::vector<T> d_data;
    bool d_repeat;
    size_t d_offset;
    size_t d_vlen;
    std::vector<tag_t> d_tags;
};

typedef vector_source<std::uint8_t> vector_source_b;
typedef vector_source<std::int16_t> vector_source_s;
typedef vector_source<std::int32_t> vector_source_i;
typedef vector_source<float> vector_source_f;
typedef vector_source<gr_complex> vector_source_c;

} // namespace blocks
} // namespace gr
