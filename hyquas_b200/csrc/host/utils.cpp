// Per-process globals and small numeric helpers of the host layer.
// Process model: one process drives one GPU.  With WORLD_SIZE/RANK/LOCAL_RANK in the environment
// (torchrun, mpirun-style launchers) the processes form numGPUs = WORLD_SIZE shards of the state, which is
// the reference's USE_MPI layout with one GPU per rank (src/utils.cpp:17-74); without them numGPUs = 1.
#include "utils.h"

#include <signal.h>
#include <spawn.h>
#include <sys/prctl.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iterator>
#include <vector>
#include "logger.h"
#include "swap.h"

extern char** environ;

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

// ranks started by the launcher below; a signal that ends the launcher (timeout(1), a batch system) is passed on to them
static volatile pid_t spawnedRanks[64];
static volatile int numSpawnedRanks = 0;
static void forwardSignal(int sig) {
    for (int i = 0; i < numSpawnedRanks; i++) if (spawnedRanks[i] > 0) kill(spawnedRanks[i], sig);
}

// Launcher mode.  The reference's single-process build drives every visible GPU from one `./main file.qasm`
// (src/utils.cpp:17-60); here one process drives one GPU, so a native program started WITHOUT a launcher environment
// (no RANK / WORLD_SIZE) on a box with several GPUs re-executes itself once per GPU -- RANK, LOCAL_RANK, WORLD_SIZE, MASTER_PORT
// set -- waits for the ranks and exits with their status: scripts/check_wrapper.sh style drivers work unchanged.  The GPU count is
// taken from a throw-away child so that this process never initialises CUDA.  HQ_NUM_GPUS=n picks the count (1: stay single),
// HQ_SELF_SPAWN=0 switches the mode off; interpreters (python) never self-spawn -- they are launched by torchrun.
static void selfSpawnIfNeeded() {
    if (getenv("RANK") || getenv("WORLD_SIZE") || getenv("HQ_SPAWNED")) return;
    if (const char* e = getenv("HQ_SELF_SPAWN")) if (atoi(e) == 0) return;
    char exe[4096];
    const ssize_t len = readlink("/proc/self/exe", exe, sizeof(exe) - 1);
    if (len <= 0) return;
    exe[len] = 0;
    const char* base = strrchr(exe, '/');
    if (strstr(base ? base : exe, "python")) return;
    int want = envInt("HQ_NUM_GPUS", 0);
    if (want == 1) return;
    std::vector<std::string> args;
    {
        std::ifstream in("/proc/self/cmdline", std::ios::binary);
        std::string all((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
        size_t pos = 0;
        while (pos < all.size()) { args.push_back(all.c_str() + pos); pos += args.back().size() + 1; }
    }
    if (args.empty()) return;
    auto spawn = [&](const std::vector<std::string>& extraEnv, int outFd) {
        std::vector<char*> argv;
        for (auto& a : args) argv.push_back(const_cast<char*>(a.c_str()));
        argv.push_back(nullptr);
        std::vector<std::string> envs;
        for (char** e = environ; *e; e++) envs.push_back(*e);
        for (auto& e : extraEnv) envs.push_back(e);
        std::vector<char*> envp;
        for (auto& e : envs) envp.push_back(const_cast<char*>(e.c_str()));
        envp.push_back(nullptr);
        posix_spawn_file_actions_t fa;
        posix_spawn_file_actions_init(&fa);
        if (outFd >= 0) posix_spawn_file_actions_adddup2(&fa, outFd, 1);
        pid_t pid = -1;
        const int rc = posix_spawn(&pid, exe, &fa, nullptr, argv.data(), envp.data());
        posix_spawn_file_actions_destroy(&fa);
        return rc == 0 ? pid : (pid_t)-1;
    };
    int visible = 0;
    {   // GPU count from a child (HQ_COUNT_GPUS): CUDA stays uninitialised here
        int fds[2];
        if (pipe(fds) != 0) return;
        const pid_t pid = spawn({"HQ_COUNT_GPUS=1", "HQ_SPAWNED=1"}, fds[1]);
        close(fds[1]);
        if (pid < 0) { close(fds[0]); return; }
        char buf[64] = {0};
        const ssize_t n = read(fds[0], buf, sizeof(buf) - 1);
        close(fds[0]);
        int st = 0;
        waitpid(pid, &st, 0);
        if (n > 0) visible = atoi(buf);
    }
    int n = 1;
    while (n * 2 <= visible) n *= 2;
    if (want > 1) { int w = 1; while (w * 2 <= want) w *= 2; n = std::min(n, w); }
    // A circuit file on the command line tells how many qubits there are: keep at least 20 of them local (a 2^20-amplitude shard
    // is 16 MiB; below that more GPUs only add exchanges, and below 10 local qubits compile() refuses).
    for (size_t a = 1; a < args.size() && want <= 1; a++) {
        std::ifstream f(args[a]);
        if (!f) continue;
        std::string head(4096, '\0');
        f.read(&head[0], (std::streamsize)head.size());
        const size_t q = head.find("qreg");
        const size_t lb = q == std::string::npos ? q : head.find('[', q);
        if (lb == std::string::npos) continue;
        const int qubits = atoi(head.c_str() + lb + 1);
        while (n > 1 && qubits - get_bit(n) < 20) n /= 2;
        break;
    }
    if (n <= 1) return;
    const std::string port = "MASTER_PORT=" + std::to_string(20000 + (int)(getpid() % 20000));
    const std::string run = "TORCHELASTIC_RUN_ID=hq" + std::to_string((long long)time(nullptr));
    std::vector<pid_t> kids;
    for (int sig : {SIGTERM, SIGINT, SIGHUP}) {
        struct sigaction sa;
        memset(&sa, 0, sizeof(sa));
        sa.sa_handler = forwardSignal;
        sigaction(sig, &sa, nullptr);
    }
    for (int r = 0; r < n; r++) {
        const pid_t pid = spawn({"RANK=" + std::to_string(r), "LOCAL_RANK=" + std::to_string(r), "WORLD_SIZE=" + std::to_string(n),
                                 "HQ_SPAWNED=" + std::to_string((long)getpid()), port, run, "MASTER_ADDR=127.0.0.1"}, -1);
        if (pid < 0) { fprintf(stderr, "hyquas_b200: cannot start rank %d\n", r); for (pid_t k : kids) kill(k, SIGTERM); exit(1); }
        kids.push_back(pid);
        if (numSpawnedRanks < 64) { spawnedRanks[numSpawnedRanks] = pid; numSpawnedRanks = numSpawnedRanks + 1; }
    }
    int worst = 0;
    for (pid_t k : kids) {
        int st = 0;
        while (waitpid(k, &st, 0) < 0 && errno == EINTR) {}
        const int code = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + (WIFSIGNALED(st) ? WTERMSIG(st) : 0);
        if (code != 0 && worst == 0) { worst = code; for (pid_t o : kids) if (o != k) kill(o, SIGTERM); }
    }
    fflush(stdout);
    _exit(worst);
}

void init() {
    if (getenv("HQ_COUNT_GPUS")) {   // the launcher's throw-away child: print the device count and leave
        int visible = 0;
        if (hq_device_count(&visible) != HQ_OK) visible = 0;
        printf("%d\n", visible);
        fflush(stdout);
        _exit(0);
    }
    if (getenv("HQ_SPAWNED")) {   // a rank of the launcher: do not outlive it
        prctl(PR_SET_PDEATHSIG, SIGTERM);
        const int launcher = atoi(getenv("HQ_SPAWNED"));
        if (launcher > 1 && (int)getppid() != launcher) _exit(1);   // it was gone before the line above took effect
    }
    selfSpawnIfNeeded();
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
