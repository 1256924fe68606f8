// Per-process globals and small numeric helpers of the host layer.
// Process model: one process drives one GPU.  With WORLD_SIZE/RANK/LOCAL_RANK in the environment
// (torchrun, mpirun-style launchers) the processes form numGPUs = WORLD_SIZE shards of the state, which is
// the reference's USE_MPI layout with one GPU per rank (src/utils.cpp:17-74); without them numGPUs = 1.
#include "utils.h"

#include <cstring>
#include "logger.h"
#include "swap.h"

namespace MyGlobalVars {
int numGPUs = 1;
int localGPUs = 1;
int bit = 0;
bool hostOnly = false;
bool swapAnyBit = false;

static int envInt(const char* key, int dflt) {
    const char* v = getenv(key);
    return v ? atoi(v) : dflt;
}

void init() {
    MyMPI::init();
    numGPUs = MyMPI::commSize;
    localGPUs = 1;
    bit = get_bit(numGPUs);
    int visible = 0;
    checkHq(hq_device_count(&visible));
    const int dev = envInt("LOCAL_RANK", MyMPI::rank) % (visible > 0 ? visible : 1);
    checkHq(hq_init(dev));
    char name[256];
    checkHq(hq_device_info(name, sizeof(name), nullptr, nullptr));
    Logger::add("Local GPU: %d", localGPUs);
    Logger::add("[%d] %s", MyMPI::rank, name);
    if (numGPUs > 1) {
        hyquas::commInitFromEnv();
        int any = 0;
        checkHq(hq_swap_any_position(&any));
        swapAnyBit = any != 0;
    }
}

void initForTest(int worldSize, int rank) {
    hostOnly = true;
    if (const char* e = getenv("HQ_TEST_SWAP_ANY")) swapAnyBit = atoi(e) != 0;
    numGPUs = worldSize;
    localGPUs = 1;
    bit = get_bit(worldSize);
    MyMPI::rank = rank;
    MyMPI::commSize = worldSize;
    MyMPI::commBit = bit;
}
}  // namespace MyGlobalVars

namespace MyMPI {
int rank = 0;
int commSize = 1;
int commBit = 0;
void init() {
    rank = MyGlobalVars::envInt("RANK", 0);
    commSize = MyGlobalVars::envInt("WORLD_SIZE", 1);
    commBit = get_bit(commSize);
}
}  // namespace MyMPI

qreal zero_wrapper(qreal x) {   // |x| < 1e-14 prints as +0 (dump format, reference src/utils.cpp:77-84)
    return (x > -1e-14 && x < 1e-14) ? 0 : x;
}

qComplex operator * (const qComplex& a, const qComplex& b) {
    return make_qComplex(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

qComplex operator + (const qComplex& a, const qComplex& b) { return make_qComplex(a.x + b.x, a.y + b.y); }

qComplex make_qComplex(qreal x) { return make_qComplex(x, 0.0); }

bool operator < (const qComplex& a, const qComplex& b) { return a.x == b.x ? a.y < b.y : a.x < b.x; }

int get_bit(int n) {
    int b = 0;
    while ((1 << b) < n) b++;
    if (n <= 0 || (1 << b) != n) {
        printf("Must be pow of two: %d\n", n);
        exit(1);
    }
    return b;
}
