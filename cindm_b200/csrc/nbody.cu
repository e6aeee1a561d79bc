#include "engine.h"
namespace cindm {
int nbody_rollout(const double*, double*, int, int, int, int, cudaStream_t) { return fail(-99, "nbody rollout not built yet"); }
int score_designs(const float*, float*, float*, int, int, int, double, double, cudaStream_t) { return fail(-99, "score not built yet"); }
}
