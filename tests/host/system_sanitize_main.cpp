// tests/host/system_sanitize_main.cpp -- TEST INFRASTRUCTURE ONLY: drives the host-side System state machine (over the CPU oracle
// backend) through a frame file incl. a blackout (reset + re-initialisation) and findPlane calls; built with
// -fsanitize=address,undefined by tools/sanitize_system_core.sh.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "system_cpu_backend.cpp"
int main(int argc, char** argv) {
    const int w = atoi(argv[2]), h = atoi(argv[3]), nf = atoi(argv[4]);
    FILE* f = fopen(argv[1], "rb");
    std::vector<uint8_t> frames((size_t)w * h * 4 * nf);
    if (fread(frames.data(), 1, frames.size(), f) != frames.size()) return 2;
    void* s = cpu_system_create(w, h, atof(argv[5]), atof(argv[6]), atof(argv[7]), atof(argv[8]));
    double T[7]; float pl[16]; int st = 0, planes = 0;
    std::vector<uint8_t> black((size_t)w * h * 4, 0);
    for (int k = 0; k < nf; k++) {
        st = cpu_system_process(s, frames.data() + (size_t)k * w * h * 4, k * 33.333, T);
        if (k % 10 == 9) planes += cpu_system_find_plane(s, pl, 100);
        if (k == nf / 2) { st = cpu_system_process(s, black.data(), k * 33.333 + 5, T); printf("blackout -> %d\n", st); }
    }
    printf("last status %d planes %d\n", st, planes);
    cpu_system_destroy(s);
    return 0;
}
