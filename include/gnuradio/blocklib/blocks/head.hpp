// blocks::head -- passes the first nitems items, then WORK_DONE (reference
// blocklib/blocks/include/gnuradio/blocklib/blocks/head.hpp:10-75).  Host buffers; the copy is
// a memcpy.  For device edges see gnuradio/blocklib/cuda/head.hpp.
#pragma once
#include <gnuradio/sync_block.hpp>

#include <cstring>

namespace gr {
namespace blocks {

class head : public sync_block
{
public:
    typedef std::shared_ptr<head> sptr;
    static sptr make(size_t itemsize, size_t nitems)
    {
        auto ptr = std::make_shared<head>(itemsize, nitems);
        ptr->add_port(untyped_port::make("input", port_direction_t::INPUT, itemsize));
        ptr->add_port(untyped_port::make("output", port_direction_t::OUTPUT, itemsize));
        return ptr;
    }
    head(size_t itemsize, size_t nitems) : sync_block("head"), _itemsize(itemsize), _nitems(nitems) {}
    work_return_code_t work(std::vector<block_work_input>& work_input,
                            std::vector<block_work_output>& work_output) override
    {
        if (_ncopied >= _nitems) {
            work_output[0].n_produced = 0;
            return work_return_code_t::WORK_DONE;
        }
        size_t n = std::min<size_t>(_nitems - _ncopied, (size_t)work_output[0].n_items);
        if (n)
            memcpy(work_output[0].buffer->write_ptr(), work_input[0].buffer->read_ptr(), n * _itemsize);
        _ncopied += n;
        work_output[0].n_produced = (int)n;
        return _ncopied >= _nitems ? work_return_code_t::WORK_DONE : work_return_code_t::WORK_OK;
    }

private:
    size_t _itemsize, _nitems, _ncopied = 0;
};

} // namespace blocks
} // namespace gr
