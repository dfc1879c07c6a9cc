/* TEST INFRASTRUCTURE (oracle) — stands in for the header the reference's CMake generates
 * (src/explicit/CMakeLists.txt:8-30); main.C only prints these three strings. */
#pragma once
#define GIT_COMMIT_HASH "oracle-build"
#define PROJECT_VERSION "0.0.1"
#define BUILD_DATE "n/a"
