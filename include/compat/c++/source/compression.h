/*
 * compression.h — source-compatible replacement of the reference's only native header
 * (kodak_tensorflow/lossless/c++/source/compression.h:41-45), for callers that want to keep their binding unchanged.
 *
 * The reference's Cython module (lossless/interface_cython.pyx:6-11, 54-58) says
 *     cdef extern from "c++/source/compression.h":  numpy.uint32_t compress_lossless(...) except +
 * With `-I<repo>/include/compat` in front of the include path and `-leae_b200` instead of the five C++ sources, that
 * .pyx compiles UNMODIFIED against this file: same name, same C++ signature (reference parameters included), same
 * exceptions with the same messages (compression.cpp:9-12, 32-62), the work done on the GPU by libeae_b200.so through
 * its C ABI (include/eae_b200.h: eae_compress_lossless). oracle/Makefile builds exactly that extension from the
 * reference's own .pyx and tests/test_gpu_cython_dropin.py runs it.
 */
#ifndef COMPRESSION_H
#define COMPRESSION_H

#include <cstdint>
#include <stdexcept>
#include <string>

#include "eae_b200.h"

inline uint32_t compress_lossless(uint32_t const& size,
                                  const int16_t* const array_input,
                                  int16_t* const array_output,
                                  uint8_t const& truncated_unary_length,
                                  const double* const probabilities)
{
    uint32_t nb_bits = 0;
    const int code = eae_compress_lossless(size, array_input, array_output, truncated_unary_length, probabilities, &nb_bits);
    if (code == EAE_SUCCESS) return nb_bits;
    const char* text = eae_last_error();
    const std::string message(text ? text : "");
    if (code == EAE_ERR_NULL) throw std::invalid_argument("One of the three pointers is NULL.");      // compression.cpp:9-12
    if (code == EAE_ERR_UNARY_LENGTH) throw std::out_of_range(message);                               // vector::at, LosslessCoder.cpp
    // 1..4: "Error of type N during the encoding." etc. (compression.cpp:32-62); anything else is a CUDA / argument failure
    throw std::runtime_error(message.empty() ? "Error of type " + std::to_string(code) + " during the encoding." : message);
}

#endif  // COMPRESSION_H
