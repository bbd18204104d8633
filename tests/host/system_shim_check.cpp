// tests/host/system_shim_check.cpp -- compiles include/alva_system.hpp (the reference's `class System` over the C ABI) the way
// the reference's embind.cpp would use it: member-function pointers with the reference's exact signatures must exist.
#include "../../include/alva_system.hpp"
#include <cstdio>
#include <type_traits>

int main() {
    // the signatures embind.cpp:9-18 binds (system.hpp:28-38)
    void (System::*cfg)(int, int, double, double, double, double, double, double, double, double) = &System::configure;
    void (System::*rst)() = &System::reset;
    int (System::*pose)(int, int) = &System::findCameraPose;
    int (System::*imu)(int, int, int) = &System::findCameraPoseWithIMU;
    int (System::*plane)(int, int) = &System::findPlane;
    int (System::*pts)(int) = &System::getFramePoints;
    (void)cfg; (void)rst; (void)pose; (void)imu; (void)plane; (void)pts;
    System s;
    float p[16];
    uint8_t px[4] = {0, 0, 0, 255};
    const int st = s.findCameraPose(px, p);          // not configured: ALVA_E_STATE, never a crash
    std::printf("unconfigured findCameraPose -> %d\n", st);
    return st == ALVA_E_STATE ? 0 : 1;
}
