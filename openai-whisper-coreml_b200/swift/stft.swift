//
//  stft.swift — unchanged call sequence of the reference's Whisper/Whisper/stft.swift:8-19; the C symbol now resolves
//  to libwhisper_b200 (f64 on the GPU) instead of libstft.a.
//
func generateSpectrogram(audio: [Double]) -> [Double] {
    var audio = audio
    audio.insert(contentsOf: [Double](repeating: 0, count: 200), at: 0)
    audio.append(contentsOf: [Double](repeating: 0, count: 200))
    var result = [Double](repeating: 0, count: 80 * 3000)
    audio.withUnsafeMutableBufferPointer { audioPtr in
        result.withUnsafeMutableBufferPointer { resultPtr in
            generate_spectrogram(audioPtr.baseAddress!, resultPtr.baseAddress!)
        }
    }
    return result
}
