"""WHIR parameter derivation: the host-side mirror of `WhirConfig::new` with `SecurityAssumption::JohnsonBound`
(crates/whir/src/config.rs:146-617: FoldingFactor, compute_number_of_rounds, queries / ood_samples /
folding_pow_bits and the log_c search of compute_optimal_log_c_for_rate).  Pure f64 arithmetic like the reference;
the derived schedule for the production parameters is pinned in tests/test_whir_protocol.py against the numbers
the reference documents (230/74 queries at 2^22, 256/75/32 at 2^28).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field as dc_field

from . import field as F


def two_adic_generator(bits: int) -> int:
    """Montgomery form of the generator of the 2^bits-th roots of unity (koala-bear/src/koala_bear.rs two-adic table)"""
    return F.two_adic_generator(bits) * F._R % F.P

def _log_eta(log_inv_rate, log_c):
    return -(0.5 * log_inv_rate + log_c)  # Johnson bound


def _list_size_bits(log_inv_rate, log_c):
    return log_inv_rate / 2.0 - (1.0 + _log_eta(log_inv_rate, log_c))


def _prox_gaps_error(log_degree, log_inv_rate, field_bits, num_functions, log_c):
    eta = 2.0 ** _log_eta(log_inv_rate, log_c)
    rho = 1.0 / float(1 << log_inv_rate)
    rho_sqrt = math.sqrt(rho)
    gamma = 1.0 - rho_sqrt - eta
    n = float(1 << (log_degree + log_inv_rate))
    m = max(math.ceil(rho_sqrt / (2.0 * eta)), 3.0)
    num_1 = (2.0 * (m + 0.5) ** 5 + 3.0 * (m + 0.5) * gamma * rho) * n
    den_1 = 3.0 * rho * rho_sqrt
    error = math.log2(num_1 / den_1 + (m + 0.5) / rho_sqrt)
    return field_bits - (error + math.log2(num_functions - 1.0))


def _log_1_delta(log_inv_rate, log_c):
    eta = 2.0 ** _log_eta(log_inv_rate, log_c)
    rate = 1.0 / float(1 << log_inv_rate)
    return math.log2(1.0 - (1.0 - math.sqrt(rate) - eta))


def _queries(level, log_inv_rate, log_c):
    return math.ceil(-level / _log_1_delta(log_inv_rate, log_c))


def _queries_error(log_inv_rate, nq, log_c):
    return -nq * _log_1_delta(log_inv_rate, log_c)


def _ood_samples(level, log_degree, log_inv_rate, field_bits, log_c):
    for s in range(1, 64):
        err = s * field_bits + 1.0 - (2.0 * _list_size_bits(log_inv_rate, log_c) + log_degree * s)
        if err >= level:
            return s
    raise AssertionError("Could not find an appropriate number of OOD samples")


def _folding_pow_bits(level, field_bits, num_variables, log_inv_rate, log_c):
    prox = _prox_gaps_error(num_variables, log_inv_rate, field_bits, 2, log_c)
    sumcheck = field_bits - (_list_size_bits(log_inv_rate, log_c) + 1.0)
    return max(0.0, level - min(prox, sumcheck))


@dataclass
class RoundConfig:
    query_pow_bits: int
    folding_pow_bits: int
    num_queries: int
    ood_samples: int
    log_inv_rate: int
    num_variables: int
    folding_factor: int
    domain_size: int
    folded_domain_gen: int  # Montgomery u32


@dataclass
class WhirConfig:
    """WhirConfig::new with SecurityAssumption::JohnsonBound (config.rs:186-334)."""

    num_variables: int
    security_level: int = 124
    pow_bits: int = 16
    first_folding: int = 7
    subsequent_folding: int = 5
    rs_domain_initial_reduction_factor: int = 5
    max_num_variables_to_send_coeffs: int = 8
    starting_log_inv_rate: int = 1
    field_bits: int = 155  # EF::bits(): bit length of p^5
    round_parameters: list = dc_field(default_factory=list)

    def folding_at(self, r):
        return self.first_folding if r == 0 else self.subsequent_folding

    def total_folding(self, n_rounds):
        return self.first_folding + self.subsequent_folding * n_rounds

    def rs_reduction_factor(self, r):
        return self.rs_domain_initial_reduction_factor if r == 0 else 1

    def _optimal_log_c(self, num_variables, log_inv_rate):
        level = max(self.security_level - self.pow_bits, 0)
        best_m, best_q = 3, None
        for m in range(3, 101):
            log_c = math.log2(2.0 * m)
            if math.ceil(_folding_pow_bits(self.security_level, self.field_bits, num_variables, log_inv_rate, log_c)) > self.pow_bits:
                break
            q = _queries(level, log_inv_rate, log_c)
            if best_q is None or q < best_q:
                best_q, best_m = q, m
        return math.log2(2.0 * best_m)

    def __post_init__(self):
        nv = self.num_variables
        assert 0 < self.first_folding <= nv and 0 < self.subsequent_folding <= nv
        assert self.rs_domain_initial_reduction_factor <= self.first_folding
        level = max(self.security_level - self.pow_bits, 0)
        log_inv_rate = self.starting_log_inv_rate
        domain_size = 1 << (nv + log_inv_rate)
        assert nv + log_inv_rate - self.first_folding <= 24, "Increase folding_factor_0"
        rest = nv - self.first_folding
        if rest < self.max_num_variables_to_send_coeffs:
            num_rounds, self.final_sumcheck_rounds = 0, rest
        else:
            num_rounds = -(-(rest - self.max_num_variables_to_send_coeffs) // self.subsequent_folding)
            self.final_sumcheck_rounds = rest - num_rounds * self.subsequent_folding
        log_c_old = self._optimal_log_c(nv, log_inv_rate)
        self.commitment_ood_samples = _ood_samples(self.security_level, nv, log_inv_rate, self.field_bits, log_c_old)
        self.starting_folding_pow_bits = math.ceil(
            _folding_pow_bits(self.security_level, self.field_bits, nv, log_inv_rate, log_c_old))
        self.round_parameters = []
        nvm = nv - self.first_folding
        for rnd in range(num_rounds):
            rs_red = self.rs_reduction_factor(rnd)
            next_rate = log_inv_rate + (self.folding_at(rnd) - rs_red)
            log_c_new = self._optimal_log_c(nvm, next_rate)
            num_queries = _queries(level, log_inv_rate, log_c_old)
            ood = _ood_samples(self.security_level, nvm, next_rate, self.field_bits, log_c_new)
            query_error = _queries_error(log_inv_rate, num_queries, log_c_old)
            comb_error = self.field_bits - (math.log2(ood + num_queries) + _list_size_bits(next_rate, log_c_new) + 1.0)
            query_pow = max(0.0, self.security_level - min(query_error, comb_error))
            fold_pow = _folding_pow_bits(self.security_level, self.field_bits, nvm, next_rate, log_c_new)
            ff = self.folding_at(rnd)
            gen = two_adic_generator(domain_size.bit_length() - 1 - ff)
            self.round_parameters.append(RoundConfig(math.ceil(query_pow), math.ceil(fold_pow), num_queries, ood, log_inv_rate,
                                                     nvm, ff, domain_size, gen))
            nvm -= self.folding_at(rnd + 1)
            log_inv_rate = next_rate
            domain_size >>= rs_red
            log_c_old = log_c_new
        self.final_queries = _queries(level, log_inv_rate, log_c_old)
        self.final_query_pow_bits = math.ceil(max(0.0, self.security_level - _queries_error(log_inv_rate, self.final_queries, log_c_old)))
        self.final_log_inv_rate = log_inv_rate

    @property
    def n_rounds(self):
        return len(self.round_parameters)

    def starting_domain_size(self):
        return 1 << (self.num_variables + self.starting_log_inv_rate)

    def n_vars_of_final_polynomial(self):
        return self.num_variables - self.total_folding(self.n_rounds)

    def final_round_config(self):
        assert self.round_parameters, "no WHIR round (config.rs:422)"
        last = self.round_parameters[-1]
        rs_red = self.rs_reduction_factor(self.n_rounds - 1)
        ff = self.folding_at(self.n_rounds)
        domain_size = last.domain_size >> rs_red
        return RoundConfig(self.final_query_pow_bits, 0, self.final_queries, last.ood_samples, last.log_inv_rate,
                           last.num_variables - ff, ff, domain_size, two_adic_generator(domain_size.bit_length() - 1 - ff))
