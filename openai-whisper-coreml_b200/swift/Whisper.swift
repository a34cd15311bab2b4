//
//  Whisper.swift — the reference's `struct Whisper` (Whisper/Whisper/Whisper.swift:11-41) re-pointed at libwhisper_b200.
//  Same surface: init() throws, encode(audio:) , decode(audioFeatures:) printing a language code, LANGUAGES.
//  The CoreML `encoder`/`decoder` model classes are replaced by C-ABI calls; MLMultiArray by [Float].
//  Shipped uncompiled (no Swift toolchain in the build image); tests/test_gpu_parity.py drives the identical sequence.
//
import Foundation

enum WhisperB200Error: Error { case failed(Int32, String) }

@inline(__always) func wbCheck(_ rc: Int32) throws {
    if rc != 0 { throw WhisperB200Error.failed(rc, String(cString: wb_last_error())) }
}

struct Whisper {
    static let LANGUAGES = ["en", "zh", "de", "es", "ru", "ko", "fr", "ja", "pt", "tr", "pl", "ca", "nl", "ar", "sv", "it", "id", "hi", "fi", "vi", "iw", "uk", "el", "ms", "cs", "ro", "da", "hu", "ta", "no", "th", "ur", "hr", "bg", "lt", "la", "mi", "ml", "cy", "sk", "te", "fa", "lv", "bn", "sr", "az", "sl", "kn", "et", "mk", "br", "eu", "is", "hy", "ne", "mn", "bs", "kk", "sq", "sw", "gl", "mr", "pa", "si", "km", "sn", "yo", "so", "af", "oc", "ka", "be", "tg", "sd", "gu", "am", "yi", "lo", "uz", "fo", "ht", "ps", "tk", "nn", "mt", "sa", "lb", "my", "bo", "tl", "mg", "as", "tt", "haw", "ln", "ha", "ba", "jw", "su"]

    let handle: OpaquePointer
    let dims: wb_dims

    // reference: loads encoder.mlpackage / decoder.mlpackage ("small", whisper_to_cml.py:7)
    init(weights: [String: [Float]], device: Int32 = 0) throws {
        var d = wb_dims(n_mels: 80, n_audio_ctx: 1500, n_audio_state: 768, n_audio_head: 12, n_audio_layer: 12,
                        n_vocab: 51865, n_text_ctx: 448, n_text_state: 768, n_text_head: 12, n_text_layer: 12)
        var h: OpaquePointer?
        try wbCheck(wb_create(&d, 1, 1, device, nil, &h))
        handle = h!
        dims = d
        for (name, tensor) in weights {
            try tensor.withUnsafeBufferPointer { try wbCheck(wb_set_weight(handle, name, $0.baseAddress, $0.count)) }
        }
        try wbCheck(wb_weights_commit(handle))
    }

    // reference Whisper.swift:23-31 — log-mel (generateSpectrogram) + encoderModel.prediction(x_1:)
    func encode(audio: [Double]) throws -> [Float] {
        let pcm = audio.map { Float($0) }
        var xa = [Float](repeating: 0, count: 1500 * Int(dims.n_audio_state))
        try pcm.withUnsafeBufferPointer { a in
            try xa.withUnsafeMutableBufferPointer { o in try wbCheck(wb_encode(handle, a.baseAddress, 1, o.baseAddress)) }
        }
        return xa
    }

    // reference Whisper.swift:33-40 — one decoder call on [50258], arg-max over logits 50259...50357
    func decode(audioFeatures: [Float]) throws {
        try audioFeatures.withUnsafeBufferPointer { try wbCheck(wb_set_audio_features(handle, $0.baseAddress, 1)) }
        var langIdx: Int32 = 0
        try wbCheck(wb_detect_language(handle, 1, 50258, 50259, &langIdx))
        print(Self.LANGUAGES[Int(langIdx)])
    }
}
