// gr::cuda::fuse_adjacent(flowgraph) -- "the elementwise blocks are fused where they are adjacent
// in a flowgraph": a graph-rewriting pass run before validate() that folds
//     fir_filter_ccf  -> multiply_const_cc          into the FIR epilogue,
//     multiply_const_cc -> fft                       into the FFT window / twiddle tables,
//     fft             -> complex_to_mag(_squared)    into the FFT epilogue,
// whenever the edge between the two blocks has exactly one reader (so nobody else observes the
// intermediate stream).  Every fused pair removes one full pass over HBM (16 B / sample) and one
// kernel launch per work() call; results stay within the 1e-5 tolerance of the unfused chain
// (bit-identical for the FIR epilogue).  The reference has no such pass -- its scheduler only
// groups blocks into threads (schedulers/mt/include/gnuradio/schedulers/mt/block_group_properties.hpp).
#pragma once
#include <gnuradio/blocklib/cuda/complex_to_mag.hpp>
#include <gnuradio/blocklib/cuda/fft.hpp>
#include <gnuradio/blocklib/cuda/fir_filter.hpp>
#include <gnuradio/blocklib/cuda/multiply_const.hpp>
#include <gnuradio/graph.hpp>

namespace gr {
namespace cuda {

namespace detail {

// re-point every edge that touches `from` (on the given side) to `to`, keeping custom buffers
inline void move_edges(graph& g, node_sptr from, node_sptr to, bool inputs, bool outputs)
{
    if (inputs)
        for (auto& e : g.in_edges(from)) {
            auto src = e->src();
            auto fac = e->buffer_factory();
            auto props = e->buf_properties();
            g.disconnect(e);
            auto ne = g.connect(src, node_endpoint(to, to->get_port(0, port_type_t::STREAM, port_direction_t::INPUT)));
            if (fac)
                ne->set_custom_buffer(fac, props);
        }
    if (outputs)
        for (auto& e : g.out_edges(from)) {
            auto dst = e->dst();
            auto fac = e->buffer_factory();
            auto props = e->buf_properties();
            g.disconnect(e);
            auto ne = g.connect(node_endpoint(to, to->get_port(0, port_type_t::STREAM, port_direction_t::OUTPUT)), dst);
            if (fac)
                ne->set_custom_buffer(fac, props);
        }
}

} // namespace detail

// returns the number of block pairs fused
inline int fuse_adjacent(graph& g)
{
    int fused = 0;
    bool again = true;
    while (again) {
        again = false;
        for (auto& e : g.edges()) {
            auto a = e->src().node();
            auto b = e->dst().node();
            if (g.out_edges(a).size() != 1) // somebody else reads the intermediate stream
                continue;

            // ---- fir_filter_ccf -> multiply_const_cc
            auto fir = std::dynamic_pointer_cast<fir_filter_ccf>(a);
            auto mul = std::dynamic_pointer_cast<multiply_const_cc>(b);
            if (fir && mul && mul->vlen() == 1 && !fir->has_fused_multiply_const()) {
                fir->set_fused_multiply_const(mul->k());
                g.disconnect(e);
                detail::move_edges(g, mul, fir, false, true);
                fused++;
                again = true;
                break;
            }
            // ---- multiply_const_cc -> fft
            auto mul_a = std::dynamic_pointer_cast<multiply_const_cc>(a);
            auto fft_b = std::dynamic_pointer_cast<fft>(b);
            if (mul_a && fft_b && !fft_b->fused_pre_multiply_const() &&
                ((fft_b->stream_input() && mul_a->vlen() == 1) ||
                 (!fft_b->stream_input() && mul_a->vlen() == fft_b->fft_size()))) {
                auto nf = fft::make(fft_b->fft_size(), fft_b->forward(), fft_b->window(), fft_b->shift(),
                                    fft_b->output(), fft_b->stream_input(), true, mul_a->k());
                g.disconnect(e);
                detail::move_edges(g, mul_a, nf, true, false);
                detail::move_edges(g, fft_b, nf, false, true);
                fused++;
                again = true;
                break;
            }
            // ---- fft -> complex_to_mag
            auto fft_a = std::dynamic_pointer_cast<fft>(a);
            auto mag = std::dynamic_pointer_cast<complex_to_mag>(b);
            if (fft_a && mag && fft_a->output() == fft_output_t::COMPLEX && mag->vlen() == fft_a->fft_size()) {
                auto nf = fft::make(fft_a->fft_size(), fft_a->forward(), fft_a->window(), fft_a->shift(),
                                    mag->squared() ? fft_output_t::MAG_SQUARED : fft_output_t::MAG,
                                    fft_a->stream_input(), fft_a->fused_pre_multiply_const(),
                                    fft_a->pre_multiply_const());
                g.disconnect(e);
                detail::move_edges(g, fft_a, nf, true, false);
                detail::move_edges(g, mag, nf, false, true);
                fused++;
                again = true;
                break;
            }
        }
    }
    return fused;
}

inline int fuse_adjacent(const std::shared_ptr<graph>& g) { return fuse_adjacent(*g); }

} // namespace cuda
} // namespace gr
