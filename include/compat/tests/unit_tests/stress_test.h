/* Stand-in for the reference's tests/unit_tests/stress_test.h (:1-21), which declares the two internals its stress
 * test calls: both are declared, with these signatures, by bpe_b200.h. */
#pragma once
#include "../../../bpe_b200.h"
