// gr::cuda::stream_to_vector / gr::cuda::vector_to_stream -- item-size adapters between a sample
// stream and vlen-item vectors on device edges (SURVEY.md 8(f) row 4, 7.3 "stream->vector
// item-size change").  The reference's edge takes its item size from the source port
// (runtime/lib/edge.cpp:39) and graph::connect does not check it (graph.cpp:8), so a block that
// re-labels N items of size s as one item of size N*s has to sit between e.g. fir_filter (8 B
// items) and a vector-input fft (8*N B items).  Where the consumer can take the stream itself the
// adapter is not needed at all (fft::make(..., stream_input = true) is the zero-copy form); these
// blocks exist for consumers that cannot, and move the data with one vectorised device copy per
// work() on the block's own stream.  Rate-changing: derive gr::block, set n_consumed / n_produced
// (block_work_io.hpp:21,36).
#pragma once
#include <gnuradio/blocklib/cuda/cuda_block.hpp>

namespace gr {
namespace cuda {

class stream_to_vector : public block, public stream_owner
{
public:
    typedef std::shared_ptr<stream_to_vector> sptr;
    static sptr make(size_t itemsize, size_t vlen)
    {
        auto ptr = std::make_shared<stream_to_vector>(itemsize, vlen);
        ptr->add_port(untyped_port::make("input", port_direction_t::INPUT, itemsize));
        ptr->add_port(untyped_port::make("output", port_direction_t::OUTPUT, itemsize * vlen));
        return ptr;
    }
    stream_to_vector(size_t itemsize, size_t vlen) : block("stream_to_vector"), d_itemsize(itemsize), d_vlen(vlen) {}

    work_return_code_t work(std::vector<block_work_input>& work_input,
                            std::vector<block_work_output>& work_output) override
    {
        const int64_t nv = std::min<int64_t>(work_input[0].n_items / (int64_t)d_vlen, work_output[0].n_items);
        if (nv > 0) {
            work_guard g(work_input, work_output, d_stream);
            check(b200_copy(work_output[0].buffer->write_ptr(), work_input[0].buffer->read_ptr(),
                            (size_t)nv * d_vlen * d_itemsize, d_stream),
                  "cuda::stream_to_vector");
        }
        work_input[0].n_consumed = (int)(nv * (int64_t)d_vlen);
        work_output[0].n_produced = (int)nv;
        return work_return_code_t::WORK_OK;
    }
    bool done() override
    {
        synchronize();
        return block::done();
    }

private:
    size_t d_itemsize, d_vlen;
};

class vector_to_stream : public block, public stream_owner
{
public:
    typedef std::shared_ptr<vector_to_stream> sptr;
    static sptr make(size_t itemsize, size_t vlen)
    {
        auto ptr = std::make_shared<vector_to_stream>(itemsize, vlen);
        ptr->add_port(untyped_port::make("input", port_direction_t::INPUT, itemsize * vlen));
        ptr->add_port(untyped_port::make("output", port_direction_t::OUTPUT, itemsize));
        return ptr;
    }
    vector_to_stream(size_t itemsize, size_t vlen) : block("vector_to_stream"), d_itemsize(itemsize), d_vlen(vlen) {}

    work_return_code_t work(std::vector<block_work_input>& work_input,
                            std::vector<block_work_output>& work_output) override
    {
        const int64_t nv = std::min<int64_t>(work_input[0].n_items, work_output[0].n_items / (int64_t)d_vlen);
        if (nv > 0) {
            work_guard g(work_input, work_output, d_stream);
            check(b200_copy(work_output[0].buffer->write_ptr(), work_input[0].buffer->read_ptr(),
                            (size_t)nv * d_vlen * d_itemsize, d_stream),
                  "cuda::vector_to_stream");
        }
        work_input[0].n_consumed = (int)nv;
        work_output[0].n_produced = (int)(nv * (int64_t)d_vlen);
        return work_return_code_t::WORK_OK;
    }
    bool done() override
    {
        synchronize();
        return block::done();
    }

private:
    size_t d_itemsize, d_vlen;
};

} // namespace cuda
} // namespace gr
