// Device-only carry-chain Montgomery multiplication (PTX mad.lo.cc / madc.hi.cc).
// Defines ZKB_FIELD_PTX and fe_mont_mul<N> when enabled; otherwise field.cuh's portable CIOS is used.
#pragma once
#include "field.cuh"
