// Stand-in for lodepng (not in this image): PNG textures/hfields are outside the checked path.
#ifndef ORACLE_STUB_LODEPNG_H_
#define ORACLE_STUB_LODEPNG_H_
#include <cstddef>
#include <vector>
typedef enum LodePNGColorType { LCT_GREY = 0, LCT_RGB = 2, LCT_PALETTE = 3, LCT_GREY_ALPHA = 4, LCT_RGBA = 6 } LodePNGColorType;
inline const char* lodepng_error_text(unsigned) { return "PNG decoding is not available in the oracle build"; }
namespace lodepng {
inline unsigned decode(std::vector<unsigned char>&, unsigned& w, unsigned& h, const unsigned char*, size_t,
                       LodePNGColorType = LCT_RGBA, unsigned = 8) { w = h = 0; return 1; }
}
#endif
