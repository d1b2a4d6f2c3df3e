// Command line of the encoder, same usage as the reference (src/main.cpp): jpgEnc <in.ppm> [out.jpg]
#include <iostream>
#include <string>

#include "include/Image.hpp"

int main(int argc, char* argv[]) {
    if (argc < 2) {
        std::cout << "No filename was written" << std::endl;
        return 0;
    }
    const std::string out = argc < 3 ? "noname.jpg" : argv[2];
    try {
        encodePPMFile(argv[1], out);      // == loadPPM(argv[1]).writeJPEG(out), streamed (include/Image.hpp)
    } catch (const std::exception& e) {
        std::cerr << e.what() << std::endl;
        return 1;
    }
    return 0;
}
