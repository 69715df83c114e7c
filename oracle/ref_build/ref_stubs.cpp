/*
 * ref_stubs.cpp -- TEST INFRASTRUCTURE ONLY (oracle/_ref build).
 *
 * binary_descriptor_custom.cpp is compiled by line ranges (see Makefile): the LBD *compute* path
 * (:42-188, :206-259, :302-412, :524-687, :1026-1372) is the reference's own text; its EDLines *detection* half
 * (:415-521 detect, :689-1024 OctaveKeyLines, :1374-2751 EDLineDetector) is not on SPL-SLAM's path
 * (Lineextractor.cc uses LSDDetectorC or FLD for detection and BinaryDescriptor only for compute()) and needs cv::Mat_
 * expression templates, so the members that the vtable / constructors still reference are defined here as loud stubs.
 */
#include "precomp_custom.hpp"

namespace cv {
namespace line_descriptor {

static void off_path(const char* what)
{
    throw std::runtime_error(std::string("oracle/_ref: ") + what + " is outside the compiled line ranges (EDLines detection half)");
}

void BinaryDescriptor::Params::read(const cv::FileNode&) { off_path("BinaryDescriptor::Params::read"); }
void BinaryDescriptor::Params::write(cv::FileStorage&) const { off_path("BinaryDescriptor::Params::write"); }
void BinaryDescriptor::operator()(InputArray, InputArray, std::vector<KeyLine>&, OutputArray, bool, bool) const
{ off_path("BinaryDescriptor::operator()"); }
void BinaryDescriptor::detectImpl(const Mat&, std::vector<KeyLine>&, const Mat&) const { off_path("BinaryDescriptor::detectImpl"); }
int BinaryDescriptor::OctaveKeyLines(cv::Mat&, ScaleLines&) { off_path("BinaryDescriptor::OctaveKeyLines"); return -1; }

/* the BinaryDescriptor constructor creates one EDLineDetector per octave (:224-225); only its scalar parameters are kept */
BinaryDescriptor::EDLineDetector::EDLineDetector()
{
    ksize_ = 15; sigma_ = 30.0; gradienThreshold_ = 80; anchorThreshold_ = 8; scanIntervals_ = 2; minLineLen_ = 15;
    lineFitErrThreshold_ = 1.6; bValidate_ = true;
    pFirstPartEdgeX_ = pFirstPartEdgeY_ = pFirstPartEdgeS_ = NULL;
    pSecondPartEdgeX_ = pSecondPartEdgeY_ = pSecondPartEdgeS_ = NULL;
    pAnchorX_ = pAnchorY_ = NULL;
}
BinaryDescriptor::EDLineDetector::~EDLineDetector() {}

} // namespace line_descriptor
} // namespace cv
