// granges_host.cpp -- libgranges_host.so: the CPU-side pieces of the CLI path that tests want to reach
// without a GPU (text formatting, src/data/mod.rs:49-65).  No CUDA here.
#include "format.hpp"

extern "C" {

// f64::to_string() of the reference; returns the length written (buf must hold 400 bytes).
int grh_format_f64(double v, char *buf) { return (int)(grh::format_f64(buf, v) - buf); }
}
