// Host-side mirror of the reference's interface (qmf/Types.h:24): everything is FP64.
#pragma once
#include <cstddef>
#include <cstdint>

namespace qmf {
using Double = double;
}
