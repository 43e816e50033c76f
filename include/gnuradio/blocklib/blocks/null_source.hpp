// blocks::null_source -- endless zeros (reference blocklib/blocks/include/gnuradio/blocklib/
// blocks/null_source.hpp:9-57: memset of every output window).  Host buffers only; for a
// device-resident source see gnuradio/blocklib/cuda/null_source.hpp.
#pragma once
#include <gnuradio/sync_block.hpp>

#include <cstring>

namespace gr {
namespace blocks {

class null_source : public sync_block
{
public:
    typedef std::shared_ptr<null_source> sptr;
    static sptr make(size_t itemsize, size_t nports = 1)
    {
        auto ptr = std::make_shared<null_source>(itemsize, nports);
        for (size_t i = 0; i < nports; i++)
            ptr->add_port(untyped_port::make("out" + std::to_string(i), port_direction_t::OUTPUT, itemsize));
        return ptr;
    }
    null_source(size_t itemsize, size_t nports) : sync_block("null_source"), _itemsize(itemsize), _nports(nports) {}
    work_return_code_t work(std::vector<block_work_input>&, std::vector<block_work_output>& work_output) override
    {
        for (auto& w : work_output) {
            memset(w.buffer->write_ptr(), 0, (size_t)w.n_items * _itemsize);
            w.n_produced = w.n_items;
        }
        return work_return_code_t::WORK_OK;
    }

private:
    size_t _itemsize, _nports;
};

} // namespace blocks
} // namespace gr
