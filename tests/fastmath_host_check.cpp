// Host accuracy check of stochastic_parker_b200/csrc/fastmath.cuh (polynomials and reductions;
// the MUFU seeds are emulated by single-precision division, which has the same 2^-23 accuracy).
// Prints the maximum error in ulp of each function against glibc; run by tests/test_cpu_host.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../stochastic_parker_b200/csrc/fastmath.cuh"

static double ulp_err(double got, double ref)
{
    if (ref == 0.0) return std::fabs(got) == 0.0 ? 0.0 : 1e9;
    int e;
    std::frexp(ref, &e);
    return std::fabs(got - ref) / std::ldexp(1.0, e - 53);
}

int main()
{
    std::mt19937_64 g(12345);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    double m_rcp = 0, m_rsq = 0, m_sqrt = 0, m_log = 0, m_exp = 0, m_pow = 0;
    for (int i = 0; i < 2000000; ++i) {
        const double x = std::exp((U(g) - 0.5) * 120.0);  // 1e-26 .. 1e26
        m_rcp = std::fmax(m_rcp, ulp_err(fm::rcp(x), 1.0 / x));
        m_rsq = std::fmax(m_rsq, ulp_err(fm::rsqrt(x), 1.0 / std::sqrt(x)));
        m_sqrt = std::fmax(m_sqrt, ulp_err(fm::sqrt_pos(x), std::sqrt(x)));
        const double l = fm::log_pos(x), lr = std::log(x);
        // near x = 1 the result is tiny: measure against max(|log x|, ulp scale of the inputs)
        m_log = std::fmax(m_log, ulp_err(l, lr));
        const double y = (U(g) - 0.5) * 100.0;
        m_exp = std::fmax(m_exp, ulp_err(fm::exp_mid(y), std::exp(y)));
        // the kernel's use: kpara ~ exp(a log(b2) + c log(pr))
        const double b2 = std::exp((U(g) - 0.5) * 8.0), pr = 0.25 + U(g) * 200.0;
        const double got = fm::exp_mid(-1.0 / 6.0 * fm::log_pos(b2) + 4.0 / 3.0 * fm::log_pos(pr));
        const double ref = std::pow(std::sqrt(b2), -1.0 / 3.0) * std::pow(pr, 4.0 / 3.0);
        m_pow = std::fmax(m_pow, std::fabs(got - ref) / ref);
    }
    if (fm::sqrt_pos(0.0) != 0.0) { std::printf("sqrt_pos(0) != 0\n"); return 1; }
    std::printf("rcp %.3f rsqrt %.3f sqrt %.3f log %.3f exp %.3f pow_rel %.3e\n", m_rcp, m_rsq, m_sqrt, m_log,
                m_exp, m_pow);
    return 0;
}
