// `fasim` — drop-in command line of the reference (README.md:56-64); all work happens in libfasim_b200.so.
#include "../../include/fasim_b200.h"

int main(int argc, char** argv) { return ltg_main(argc, argv); }
