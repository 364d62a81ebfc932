// A subscriber that only counts: plays the decoders (JAERO) behind the PUB sockets of the publish leg of bench.py, in its own
// process and one thread per address, so that the receiving side of the measurement is not a Python loop. Test / bench tooling,
// not part of the library. libzmq has no headers in this image: the entry points are resolved at run time like publisher.cpp does.
//   zmq_sink <address> [<address> ...]      reads stdin until EOF, then prints {"messages": N, "payload_bytes": B}
#include <dlfcn.h>
#include <glob.h>
#include <unistd.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

namespace {
struct Api {
    void *(*ctx_new)();
    void *(*socket)(void *, int);
    int (*close)(void *);
    int (*setsockopt)(void *, int, const void *, size_t);
    int (*getsockopt)(void *, int, void *, size_t *);
    int (*connect)(void *, const char *);
    int (*recv)(void *, void *, size_t, int);
} z;
constexpr int kSUB = 2, kSUBSCRIBE = 6, kRCVMORE = 13, kRCVHWM = 24, kRCVTIMEO = 27, kLINGER = 17;

bool load() {
    std::vector<std::string> cands;
    if (const char *e = getenv("SDRB_LIBZMQ")) cands.push_back(e);
    cands.push_back("libzmq.so.5");
    cands.push_back("libzmq.so");
    const char *pats[] = {"/opt/prime-rl/.venv/lib/python3*/site-packages/pyzmq.libs/libzmq-*.so*",
                          "/usr/lib/python3/dist-packages/pyzmq.libs/libzmq-*.so*",
                          "/usr/local/lib/python3*/site-packages/pyzmq.libs/libzmq-*.so*"};
    for (const char *pat : pats) {
        glob_t g;
        if (glob(pat, 0, nullptr, &g) == 0) {
            if (g.gl_pathc > 0) cands.push_back(g.gl_pathv[0]);
            globfree(&g);
        }
    }
    void *lib = nullptr;
    for (const std::string &c : cands)
        if ((lib = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL))) break;
    if (!lib) return false;
    z.ctx_new = (decltype(z.ctx_new))dlsym(lib, "zmq_ctx_new");
    z.socket = (decltype(z.socket))dlsym(lib, "zmq_socket");
    z.close = (decltype(z.close))dlsym(lib, "zmq_close");
    z.setsockopt = (decltype(z.setsockopt))dlsym(lib, "zmq_setsockopt");
    z.getsockopt = (decltype(z.getsockopt))dlsym(lib, "zmq_getsockopt");
    z.connect = (decltype(z.connect))dlsym(lib, "zmq_connect");
    z.recv = (decltype(z.recv))dlsym(lib, "zmq_recv");
    return z.ctx_new && z.socket && z.setsockopt && z.getsockopt && z.connect && z.recv;
}
}  // namespace

int main(int argc, char **argv) {
    if (argc < 2 || !load()) { fprintf(stderr, "usage: zmq_sink <address>...  (libzmq must be loadable)\n"); return 2; }
    void *ctx = z.ctx_new();
    std::atomic<bool> stop{false};
    std::atomic<unsigned long long> messages{0}, bytes{0};
    std::vector<std::thread> th;
    for (int a = 1; a < argc; ++a) {
        void *s = z.socket(ctx, kSUB);
        const int zero = 0, tmo = 50;
        z.setsockopt(s, kRCVHWM, &zero, sizeof zero);
        z.setsockopt(s, kRCVTIMEO, &tmo, sizeof tmo);
        z.setsockopt(s, kLINGER, &zero, sizeof zero);
        z.setsockopt(s, kSUBSCRIBE, "", 0);
        if (z.connect(s, argv[a]) != 0) { fprintf(stderr, "zmq_sink: cannot connect to %s\n", argv[a]); return 3; }
        th.emplace_back([s, &stop, &messages, &bytes] {
            std::vector<unsigned char> buf(1 << 20);
            unsigned long long m = 0, b = 0;
            int frame = 0;
            while (!stop.load(std::memory_order_relaxed)) {
                const int n = z.recv(s, buf.data(), buf.size(), 0);
                if (n < 0) continue;                              // timeout
                int more = 0;
                size_t len = sizeof more;
                z.getsockopt(s, kRCVMORE, &more, &len);
                if (frame == 2) b += (unsigned long long)n;       // the payload frame of the reference's three
                if (more) ++frame;
                else { ++m; frame = 0; }
            }
            messages += m;
            bytes += b;
            z.close(s);
        });
    }
    printf("ready\n");
    fflush(stdout);
    char c;
    while (read(0, &c, 1) > 0) {}
    usleep(300000);                                               // what is still in flight
    stop = true;
    for (std::thread &t : th) t.join();
    printf("{\"messages\": %llu, \"payload_bytes\": %llu}\n", messages.load(), bytes.load());
    return 0;
}
