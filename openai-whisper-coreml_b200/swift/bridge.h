//
//  bridge.h — Swift bridging header for libwhisper_b200 (replaces Whisper/Whisper/bridge.h of the reference).
//  Shipped as text: there is no swiftc in the build image; the same call sequence is exercised through ctypes in tests/.
//
#ifndef bridge_h
#define bridge_h
#include "whisper_b200.h"   // declares generate_spectrogram(double*, double*) with the reference's contract, plus wb_*
#endif /* bridge_h */
