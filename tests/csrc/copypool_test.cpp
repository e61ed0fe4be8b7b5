// CopyPool (csrc/hostcopy.cuh) on the host alone: every byte arrives, nothing outside [dst, dst + n) is touched, bursts of
// copies and idle gaps (workers asleep) both work, the pool shuts down cleanly, and a forked child can still copy and exit.
// Built by tests/test_abi_host.py with g++ (and once more with -fsanitize=thread where the toolchain has it).
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <vector>

#include "hostcopy.cuh"

using l3d::CopyPool;

static int check_copy(CopyPool& pool, size_t bytes, unsigned seed) {
    std::vector<unsigned char> a(bytes + 64), b(bytes + 64, 0xAB);
    for (size_t i = 0; i < a.size(); i++) a[i] = (unsigned char)((i * 2654435761u + seed) >> 11);
    pool.copy(b.data() + 5, a.data() + 3, bytes);
    if (memcmp(b.data() + 5, a.data() + 3, bytes) != 0) return 1;
    for (int i = 0; i < 5; i++) if (b[i] != 0xAB) return 2;
    for (size_t i = bytes + 5; i < b.size(); i++) if (b[i] != 0xAB) return 3;
    return 0;
}

int main() {
    const size_t sizes[] = {0, 1, 4095, 4096, 4097, 511 << 10, 512 << 10, (512 << 10) + 1, 2764800, 3686400, (8 << 20) + 13};
    for (int nt : {1, 2, 3, 4, 8}) {
        CopyPool pool(nt);
        for (int round = 0; round < 3; round++) {
            for (size_t n : sizes) {
                int rc = check_copy(pool, n, (unsigned)(nt * 131 + round));
                if (rc) { printf("FAIL threads %d bytes %zu rc %d\n", nt, n, rc); return 1; }
            }
            usleep(round == 1 ? 20000 : 0);   // let the workers fall asleep between bursts
        }
    }
    {   // fork with live worker threads: the child copies single-threaded and destroys its copy of the pool without joining
        CopyPool pool(4);
        if (check_copy(pool, 4 << 20, 7)) { printf("FAIL before fork\n"); return 1; }
        pid_t pid = fork();
        if (pid == 0) {
            int rc = check_copy(pool, 4 << 20, 9);
            pool.~CopyPool();
            _exit(rc ? 40 + rc : 0);
        }
        int st = 0;
        waitpid(pid, &st, 0);
        if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) { printf("FAIL in forked child: status %d\n", st); return 1; }
        if (check_copy(pool, 4 << 20, 11)) { printf("FAIL after fork\n"); return 1; }
    }
    printf("copypool ok\n");
    return 0;
}
