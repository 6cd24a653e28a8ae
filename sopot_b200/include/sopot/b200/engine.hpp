// sopot/b200/engine.hpp -- RAII C++ face of the C ABI (include/sopot_b200.h).
// Error convention of the reference is preserved: bad arguments -> std::invalid_argument, everything
// else -> std::runtime_error (the reference throws both from constructors / the RHS, e.g.
// physics/harmonic_oscillator.hpp:59-61, physics/control/cart_n_pendulum.hpp:331-333).
// There is no host integration path behind these headers: without libsopot_b200.so + a B200 every
// solve/linearize call throws.
#pragma once

#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../../include/sopot_b200.h"

namespace sopot::b200 {

enum class FpMode : uint32_t { Contract = SOPOT_B200_FP_CONTRACT, Strict = SOPOT_B200_FP_STRICT };

class Engine {
public:
    explicit Engine(int device = 0, const void* nccl_unique_id = nullptr, int rank = 0, int nranks = 1) {
        const int rc = sopot_b200_init(&m_ctx, device, nccl_unique_id, rank, nranks);
        if (rc != SOPOT_B200_OK) raise(rc, sopot_b200_last_error(nullptr));
    }
    ~Engine() { if (m_ctx) sopot_b200_destroy(m_ctx); }
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;

    sopot_b200_ctx* ctx() const noexcept { return m_ctx; }
    void check(int rc) const { if (rc != SOPOT_B200_OK) raise(rc, sopot_b200_last_error(m_ctx)); }
    void sync() const { check(sopot_b200_sync(m_ctx)); }

    // process-wide default engine on device 0 (the reference API has no context argument to thread through)
    static Engine& shared() {
        static Engine e(0);
        return e;
    }

    [[noreturn]] static void raise(int rc, const char* msg) {
        const std::string text = std::string("sopot-b200: ") + (msg ? msg : "unknown error");
        if (rc == SOPOT_B200_EINVAL) throw std::invalid_argument(text);
        throw std::runtime_error(text);
    }

private:
    sopot_b200_ctx* m_ctx = nullptr;
};

// options shared by the ensemble entry points
struct EnsembleOptions {
    FpMode fp_mode = FpMode::Contract;
    size_t store_every = 0;       // 0 = final state only; k = keep every k-th step
    bool want_status = true;
};

// SoA result of an ensemble integration: state[d][i], trajectory[snapshot][d][i]
struct EnsembleResult {
    size_t n = 0, dim = 0, n_steps = 0, store_every = 0;
    std::vector<double> state;        // [dim][n]
    std::vector<double> trajectory;   // [n_steps / store_every][dim][n]
    std::vector<int32_t> status;      // per instance: SOPOT_B200_ST_*
    double final(size_t d, size_t i) const { return state[d * n + i]; }
    double at(size_t snapshot, size_t d, size_t i) const { return trajectory[(snapshot * dim + d) * n + i]; }
    size_t snapshots() const { return store_every ? n_steps / store_every : 0; }
    // reference behaviour: a failing instance is an exception (singular mass matrix etc.)
    void throwOnFailure() const {
        for (size_t i = 0; i < status.size(); ++i) {
            if (status[i] == SOPOT_B200_ST_SINGULAR)
                throw std::runtime_error("CartNPendulum: mass matrix is singular; cannot solve for accelerations.");
            if (status[i] == SOPOT_B200_ST_NONFINITE)
                throw std::runtime_error("sopot-b200: non-finite state in instance " + std::to_string(i));
        }
    }
};

inline sopot_b200_rk4_opts make_opts(const EnsembleOptions& o, uint32_t broadcast) {
    sopot_b200_rk4_opts r{};
    r.mem = SOPOT_B200_MEM_HOST;
    r.fp_mode = static_cast<uint32_t>(o.fp_mode);
    r.store_every = o.store_every;
    r.broadcast = broadcast;
    return r;
}

// per-instance parameter: one value for the whole ensemble, or one per instance
struct Param {
    std::vector<double> v;
    Param(double x) : v{x} {}
    Param(std::vector<double> xs) : v(std::move(xs)) {}
    const double* data() const { return v.data(); }
    bool broadcast(size_t n) const {
        if (v.size() == n && n != 1) return false;
        if (v.size() == 1) return n != 1 ? true : false;
        throw std::invalid_argument("sopot-b200: ensemble parameter has " + std::to_string(v.size()) +
                                    " values for " + std::to_string(n) + " instances");
    }
};

}  // namespace sopot::b200
