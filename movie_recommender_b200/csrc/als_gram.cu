// Gram-block ALS modes (algorithm 3 and 4) -- see als.cuh.
#include "als.cuh"

namespace mrb {

struct AlsProblem::GramState {};

AlsRunInfo AlsProblem::run_gram(int, double, int) {
    throw Error(kErrArgument, "als: algorithm 3/4 not built yet");
}

}  // namespace mrb
