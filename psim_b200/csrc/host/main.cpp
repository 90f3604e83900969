// `psim model1.json model2.json ...` - same command line, progress lines and result files as the reference's
// psim/src/main.cpp:10-31, with the particle loop running on a B200 through the C ABI.
// Extra, optional controls come from the environment so that the argument list stays the reference's:
//   PSIM_SEED (default: from the clock, like the reference's random_device)
//   PSIM_DEVICES (comma-separated CUDA device indices, default "0": the phonons are shared between them)
//   PSIM_STEPS_PER_LAUNCH (default: library default)
#include "../../../include/psim_host.h"

#include <chrono>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

int main(int argc, char* argv[]) {
    if (argc > 1) {
        const char* env_seed = std::getenv("PSIM_SEED");
        const char* env_dev = std::getenv("PSIM_DEVICES");
        const char* env_spl = std::getenv("PSIM_STEPS_PER_LAUNCH");
        uint64_t seed = env_seed ? std::strtoull(env_seed, nullptr, 10)
                                 : static_cast<uint64_t>(std::chrono::system_clock::now().time_since_epoch().count());
        std::vector<int> devices;
        for (const char* c = env_dev ? env_dev : "0"; *c;) {
            char* end = nullptr;
            devices.push_back(static_cast<int>(std::strtol(c, &end, 10)));
            if (end == c) { break; }
            c = (*end == ',') ? end + 1 : end;
        }
        const int spl = env_spl ? std::atoi(env_spl) : 0;
        const std::vector<std::string> filenames(argv + 1, argv + argc);
        for (const auto& filename : filenames) {
            psim_model* model = nullptr;
            if (psim_model_load(filename.c_str(), &model) != PSIM_OK) {
                std::cerr << psim_host_last_error() << '\n';
                std::cerr << "There was an error reading the data from the file at \"" << filename << "\"\n";
                continue;
            }
            const auto t0 = std::chrono::steady_clock::now();
            const int rc = psim_model_run_devices(model, devices.data(), static_cast<int>(devices.size()), seed, spl, 1, nullptr);
            const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (rc != PSIM_OK) {
                std::cerr << psim_host_last_error() << '\n';
                psim_model_free(model);
                continue;
            }
            std::cout << "Time Taken: " << secs << "[s]\n";
            if (psim_model_export(model, filename.c_str(), secs) != PSIM_OK) { std::cerr << psim_host_last_error() << '\n'; }
            psim_model_free(model);
            ++seed;
        }
    } else {
        std::cout << "Need filenames\n";
    }
    std::cout << "done\n";
    return 0;
}
