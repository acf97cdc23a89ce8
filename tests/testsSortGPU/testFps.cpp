// testFps -- glr::ApplicationFPS of include/rtr_scene.hpp against the arithmetic of the reference's
// srcOpenGL/application.cpp:345-373, restated here step by step on a replayed clock (the reference's class lives in
// a translation unit that needs GLFW / ImGui and cannot be built here).  CPU only.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <random>
#include <vector>

#include "rtr_scene.hpp"

int main() {
    std::mt19937 rng(20241018u);
    std::uniform_real_distribution<double> dt(0.0005, 0.05);
    int bad = 0, windows = 0;
    for (uint32_t every : {10u, 1u, 7u}) {
        glr::ApplicationFPS fps;
        fps._NbFramesBetweenDisplay = every;
        fps._DisplayFPS = false;
        // the restated statistics: float members, the reference's operation order
        float last = 0.f, sum = 0.f, mn = INFINITY, mx = 0.f, avgFPS = 0.f, minFPS = 0.f, maxFPS = 0.f;
        uint32_t count = 0;
        double now = 0.0;
        for (int f = 0; f < 1000; ++f) {
            now += dt(rng);
            fps.increment(now);
            const float cur = static_cast<float>(now);
            const float d = cur - last;
            last = cur;
            sum += d;
            count++;
            if (d > mx) mx = d;
            if (d < mn) mn = d;
            fps.display(nullptr);
            if (count >= every) {
                avgFPS = 1.f / (sum / static_cast<float>(count));
                minFPS = 1.f / mx;
                maxFPS = 1.f / mn;
                count = 0; sum = 0.f; mn = INFINITY; mx = 0.f;
                windows++;
            }
            if (fps.avgFPS() != avgFPS || fps.minFPS() != minFPS || fps.maxFPS() != maxFPS || fps._LastFrame != last) bad++;
            if (avgFPS != 0.f && !(minFPS <= avgFPS && avgFPS <= maxFPS)) bad++;
        }
    }
    // known answer: ten frames of exactly 1/64 s (representable) -> 64 FPS on all three lines, printed once
    glr::ApplicationFPS k;
    char text[256] = {0};
    FILE* mem = fmemopen(text, sizeof(text) - 1, "w");
    for (int f = 1; f <= 10; ++f) { k.increment(f / 64.0); k.display(mem); }
    std::fclose(mem);
    if (k.avgFPS() != 64.f || k.minFPS() != 64.f || k.maxFPS() != 64.f) bad++;
    if (std::string(text) != "avg FPS: 64.00\nmin FPS: 64.00\nmax FPS: 64.00\n\n") bad++;
    std::printf("testFps: %d windows, %d mismatches\n", windows, bad);
    return bad ? 1 : 0;
}
