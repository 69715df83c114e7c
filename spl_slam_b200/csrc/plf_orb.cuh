// plf_orb.cuh -- geometry tables shared by the ORB kernels and their host driver.
#pragma once
#include "plf_common.cuh"

#define ORB_MAX_LEVELS 16
#define ORB_EDGE 19          // EDGE_THRESHOLD, src/ORBextractor.cc:74
#define ORB_MINB 16          // EDGE_THRESHOLD - 3: FAST search border (src/ORBextractor.cc:773)
#define ORB_HALF_PATCH 15
#define ORB_CELL_W 30        // "const float W = 30", src/ORBextractor.cc:769

struct OrbLevelGeom {
    int w, h, pitch;            // level size and row pitch of the level buffers
    int nCols, nRows, wCell, hCell;
    int cellBase;               // first FAST cell id of this level in the fused grid
    int blurTileBase;           // first blur tile id of this level
    int blurTilesX;             // CTA columns of the blur: interior strip columns + one column of edge strips
    int blurF;                  // interior 4-px strips of a row (plf_strip_interior)
    int nfeat;                  // mnFeaturesPerLevel[level]
    int rawcap;                 // capacity of the raw key list per frame
    int keptcap;                // capacity of the kept list per frame
    int nodecap;                // octree node capacity
    float scale;                // mvScaleFactor[level]
    int sizeval;                // (int)(31 * scale)
    size_t frameBytes;          // bytes of one frame of this level (pitch * h)
    size_t rawOff;              // offset (in keys) of this level inside a frame's raw key block
    size_t keptOff;             // offset (in ints) of this level inside a frame's kept block
};

struct OrbGeom {
    int nlevels, iniTh, minTh;
    int totalCells, totalBlurTiles;
    int capPerFrame;            // output capacity per frame (sum of keptcap)
    size_t rawPerFrame;         // keys per frame over all levels
    size_t keptPerFrame;        // ints per frame over all levels
    int umax[ORB_HALF_PATCH + 1];
    OrbLevelGeom lv[ORB_MAX_LEVELS];
};

// device pointers of one workspace
struct OrbPtrs {
    const uint8_t* lvl[ORB_MAX_LEVELS];   // level images, frame f at lvl[l] + f*frameStride[l]
    size_t frameStride[ORB_MAX_LEVELS];
    int pitch[ORB_MAX_LEVELS];
    uint8_t* blr[ORB_MAX_LEVELS];         // blurred levels (pitch = geom pitch, frame stride = frameBytes)
    unsigned* rawkeys;                    // [frame][rawPerFrame] packed x | y<<12 | resp<<24
    int* rawcount;                        // [frame][nlevels]
    unsigned short* knode;                // [frame][rawPerFrame] octree scratch
    int* kept;                            // [frame][keptPerFrame] key indices in final list order
    int* keptcount;                       // [frame][nlevels]
    int* status;                          // [frame] overflow flags
};
