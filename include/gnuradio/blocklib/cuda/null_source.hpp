// gr::cuda::null_source / gr::cuda::head -- device-resident stand-ins for blocks::null_source and
// blocks::head (blocklib/blocks/include/gnuradio/blocklib/blocks/null_source.hpp:33-46,
// head.hpp:38-64) so that GPU flowgraphs such as BASELINE config 2
// (null_source -> fft -> complex_to_mag -> null_sink) never touch host memory.
#pragma once
#include <gnuradio/blocklib/cuda/cuda_block.hpp>

namespace gr {
namespace cuda {

// Endless zeros written on the device.  `nitems` > 0 makes it finite (null_source + head in one
// block, avoiding a pure copy stage): WORK_DONE after that many items.
class null_source : public sync_block, public stream_owner
{
public:
    typedef std::shared_ptr<null_source> sptr;
    static sptr make(size_t itemsize, uint64_t nitems = 0, bool clear_every_call = true)
    {
        auto ptr = std::make_shared<null_source>(itemsize, nitems, clear_every_call);
        ptr->add_port(untyped_port::make("out0", port_direction_t::OUTPUT, itemsize));
        return ptr;
    }
    null_source(size_t itemsize, uint64_t nitems, bool clear)
        : sync_block("null_source"), _itemsize(itemsize), _nitems(nitems), _clear(clear)
    {
    }
    work_return_code_t work(std::vector<block_work_input>& work_input,
                            std::vector<block_work_output>& work_output) override
    {
        uint64_t n = (uint64_t)work_output[0].n_items;
        if (_nitems) {
            if (_emitted >= _nitems) {
                work_output[0].n_produced = 0;
                return work_return_code_t::WORK_DONE;
            }
            n = std::min<uint64_t>(n, _nitems - _emitted);
        }
        if (_clear) {
            work_guard g(work_input, work_output, d_stream);
            check(b200_memset(work_output[0].buffer->write_ptr(), 0, (size_t)n * _itemsize, d_stream),
                  "cuda::null_source");
        }
        _emitted += n;
        work_output[0].n_produced = (int)n;
        return (_nitems && _emitted >= _nitems) ? work_return_code_t::WORK_DONE : work_return_code_t::WORK_OK;
    }
    bool done() override
    {
        synchronize();
        return sync_block::done();
    }

private:
    size_t _itemsize;
    uint64_t _nitems, _emitted = 0;
    bool _clear;
};

class head : public sync_block, public stream_owner
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
        if (n) {
            work_guard g(work_input, work_output, d_stream);
            check(b200_copy(work_output[0].buffer->write_ptr(), work_input[0].buffer->read_ptr(), n * _itemsize,
                            d_stream),
                  "cuda::head");
        }
        _ncopied += n;
        work_output[0].n_produced = (int)n;
        return _ncopied >= _nitems ? work_return_code_t::WORK_DONE : work_return_code_t::WORK_OK;
    }
    bool done() override
    {
        synchronize();
        return sync_block::done();
    }

private:
    size_t _itemsize, _nitems, _ncopied = 0;
};

} // namespace cuda
} // namespace gr
