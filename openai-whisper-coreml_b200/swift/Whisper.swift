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

    // reference: `whisper.load_model("small")` at export time (whisper_to_cml.py:7) — here the checkpoint is read at run time
    init(checkpoint path: String, maxBatch: Int32 = 1, device: Int32 = 0) throws {
        var d = wb_dims()
        try wbCheck(wb_safetensors_read_dims(path, &d))
        var h: OpaquePointer?
        try wbCheck(wb_create(&d, maxBatch, 1, device, nil, &h))
        handle = h!
        dims = d
        var loaded: Int32 = 0
        try wbCheck(wb_load_safetensors(handle, path, &loaded))
    }

    // A recording of any length (the reference stops at one 30 s window, ContentView.swift:57-62): upstream transcribe()
    func transcribe(pcm: [Float], tokenizer: OpaquePointer?, language: Int32? = nil) throws -> (tokens: [Int32], segments: [wb_segment]) {
        // multilingual token layout (Whisper.swift:35,37 pin sot = 50258, languages from 50259)
        var sotSequence: [Int32] = [50258, 50259 + (language ?? 0), 50359]
        var suppress: [Int32] = [50258, 50358, 50359, 50360, 50361, 50362], suppressBegin: [Int32] = [220, 50257]
        var lo = wb_long_opts()
        var nSeg: Int32 = 0, nTok: Int32 = 0, lang: Int32 = -1
        var segs = [wb_segment](repeating: wb_segment(), count: 64 * (pcm.count / 480000 + 2))
        var toks = [Int32](repeating: 0, count: 256 * (pcm.count / 480000 + 2))
        try sotSequence.withUnsafeMutableBufferPointer { sot in
            try suppress.withUnsafeMutableBufferPointer { sup in
                try suppressBegin.withUnsafeMutableBufferPointer { supb in
                    lo.decode.initial_tokens = UnsafePointer(sot.baseAddress); lo.decode.n_initial = 3; lo.decode.sot_index = 0
                    lo.decode.eot = 50257; lo.decode.no_speech = 50362
                    lo.decode.suppress = UnsafePointer(sup.baseAddress); lo.decode.n_suppress = Int32(sup.count)
                    lo.decode.suppress_begin = UnsafePointer(supb.baseAddress); lo.decode.n_suppress_begin = 2
                    lo.decode.timestamps = 1; lo.decode.timestamp_begin = 50364; lo.decode.no_timestamps = 50363
                    lo.decode.max_initial_timestamp_index = 50
                    lo.compression_ratio_threshold = 2.4; lo.logprob_threshold = -1.0; lo.no_speech_threshold = 0.6
                    lo.condition_on_previous_text = 1; lo.sot_prev = 50361; lo.tokenizer = tokenizer
                    lo.detect_language = language == nil ? 1 : 0; lo.lang0 = 50259
                    try withUnsafeMutablePointer(to: &lang) { lp in
                        lo.detected_language = lp
                        try wbCheck(wb_transcribe_long(handle, pcm, Int64(pcm.count), &lo, &segs, Int32(segs.count), &nSeg, &toks,
                                                       Int32(toks.count), &nTok))
                    }
                }
            }
        }
        return (Array(toks[0..<Int(nTok)]), Array(segs[0..<Int(nSeg)]))
    }
}
