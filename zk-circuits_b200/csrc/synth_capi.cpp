// C ABI of libzkb200_synth.so (include/zkb200_synth.h): the synthetic workload generator, kept OUT of the proving library.
#include "../../include/zkb200_synth.h"
#include "synth.hpp"
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>

using namespace zkb;

struct zkb_synth {
    SynthCircuit sc;
};

namespace {
thread_local std::string g_synth_error;
template <class F>
int sguarded(F&& f) {
    try {
        return f();
    } catch (const std::exception& e) { g_synth_error = e.what(); return -1; }
}
}  // namespace

extern "C" {

const char* zkb_synth_last_error(void) { return g_synth_error.c_str(); }

// ---- synthetic workloads ----
int zkb_synth_create(unsigned min_degree_bits, int zk, size_t n_poseidon, size_t n_base_sum, size_t n_arith, size_t n_const,
                     size_t num_public_inputs, uint64_t seed, zkb_synth** out) {
    return sguarded([&] {
        if (!out) throw std::invalid_argument("out is null");
        *out = nullptr;
        if (min_degree_bits > 20 || num_public_inputs > 1024) throw std::invalid_argument("bad synthetic circuit shape");
        SynthSpec sp;
        sp.min_degree_bits = min_degree_bits; sp.zk = zk != 0;
        sp.n_poseidon = n_poseidon; sp.n_base_sum = n_base_sum; sp.n_arith = n_arith; sp.n_const = n_const;
        sp.num_public_inputs = num_public_inputs; sp.seed = seed;
        auto s = std::make_unique<zkb_synth>();
        s->sc = make_synth_circuit(sp);
        *out = s.release();
        return 0;
    });
}
int zkb_synth_create_recursion(unsigned min_degree_bits, int zk, size_t n_poseidon, size_t n_base_sum, size_t n_arith, size_t n_const,
                               size_t num_public_inputs, uint64_t seed, const size_t recursion_rows[8], zkb_synth** out) {
    return sguarded([&] {
        if (!out || !recursion_rows) throw std::invalid_argument("null argument");
        *out = nullptr;
        if (min_degree_bits > 20 || num_public_inputs > 1024) throw std::invalid_argument("bad synthetic circuit shape");
        SynthSpec sp;
        sp.min_degree_bits = min_degree_bits; sp.zk = zk != 0;
        sp.n_poseidon = n_poseidon; sp.n_base_sum = n_base_sum; sp.n_arith = n_arith; sp.n_const = n_const;
        sp.num_public_inputs = num_public_inputs; sp.seed = seed;
        sp.n_arith_ext = recursion_rows[0]; sp.n_mul_ext = recursion_rows[1]; sp.n_reducing = recursion_rows[2];
        sp.n_reducing_ext = recursion_rows[3]; sp.n_random_access = recursion_rows[4]; sp.n_exp = recursion_rows[5];
        sp.n_coset = recursion_rows[6]; sp.n_mds = recursion_rows[7];
        if (!sp.recursion()) throw std::invalid_argument("recursion_rows are all zero: use zkb_synth_create");
        auto s = std::make_unique<zkb_synth>();
        s->sc = make_synth_circuit(sp);
        *out = s.release();
        return 0;
    });
}
int zkb_synth_destroy(zkb_synth* s) { delete s; return 0; }
size_t zkb_synth_num_constants(const zkb_synth* s) { return s ? s->sc.const_sigma_values.size() - 80 : 0; }
size_t zkb_synth_common_len(const zkb_synth* s) { return s ? s->sc.common.size() : 0; }
size_t zkb_synth_degree(const zkb_synth* s) { return s ? (size_t(1) << s->sc.degree_bits) : 0; }
int zkb_synth_get(const zkb_synth* s, uint8_t* common, uint64_t* const_sigma_values, uint64_t* wires, uint64_t* public_inputs) {
    return sguarded([&] {
        if (!s) throw std::invalid_argument("synth is null");
        const SynthCircuit& sc = s->sc;
        size_t n = size_t(1) << sc.degree_bits;
        if (common) std::memcpy(common, sc.common.data(), sc.common.size());
        if (const_sigma_values)
            for (size_t c = 0; c < sc.const_sigma_values.size(); ++c) std::memcpy(const_sigma_values + c * n, sc.const_sigma_values[c].data(), n * 8);
        if (wires)
            for (size_t c = 0; c < sc.wires.size(); ++c) std::memcpy(wires + c * n, sc.wires[c].data(), n * 8);
        if (public_inputs) std::memcpy(public_inputs, sc.public_inputs.data(), sc.public_inputs.size() * 8);
        return 0;
    });
}

}  // extern "C"
