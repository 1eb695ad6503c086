// UNVERIFIED (never compiled: no Swift toolchain in the build image).  Same public names as the reference's
// Sources/Smelter/ONNXGraph.swift; every method forwards to the C ABI in include/smelter_b200.h.
import CSmelterB200
import Foundation

/// TypeDefinitions.swift:1-33
public struct Shape {
    public var channels: Int, width: Int, height: Int, depth: Int
}

/// Replaces MTLContext / MTLDevice (README.md:18): a CUDA device + stream.
public final class CUDAContext {
    let handle: OpaquePointer
    public init(device: Int32 = 0) throws {
        var h: OpaquePointer?
        try ONNXGraph.check(smelter_context_create(device, nil, &h))
        handle = h!
    }
    public func synchronize() throws { try ONNXGraph.check(smelter_context_synchronize(handle)) }
    deinit { smelter_context_destroy(handle) }
}

/// Replaces MPSImage: an NCHW fp16 device buffer.
public final class DeviceImage {
    let handle: OpaquePointer
    let owned: Bool
    public init(context: CUDAContext, batch: Int32 = 1, channels: Int32, height: Int32, width: Int32) throws {
        var h: OpaquePointer?
        try ONNXGraph.check(smelter_tensor_create(context.handle, batch, channels, height, width, &h))
        handle = h!
        owned = true
    }
    init(borrowed: OpaquePointer) { handle = borrowed; owned = false }
    /// README.md:33-39 `texture(from:)` analogue.
    public func upload(_ values: [Float]) throws {
        try values.withUnsafeBufferPointer { try ONNXGraph.check(smelter_tensor_from_float(handle, nil, $0.baseAddress, $0.count)) }
    }
    /// `texture(from: CGImage)` analogue for interleaved 8-bit pixels [N][H][W][sourceChannels]: value = byte * scale[c] + bias[c].
    public func upload(pixels: [UInt8], sourceChannels: Int32, scale: [Float]? = nil, bias: [Float]? = nil) throws {
        try pixels.withUnsafeBufferPointer { px in
            try ONNXGraph.check(smelter_tensor_from_u8(handle, nil, px.baseAddress, sourceChannels, scale, bias))
        }
    }
    /// toFloatArray() in the reference's own element order (slices of four channels, MPSImage+Extensions.swift:26-59).
    public func toFloatArrayMPSOrder() -> [Float]? {
        var dims = [Int32](repeating: 0, count: 4)
        guard smelter_tensor_dims(handle, &dims) == 0 else { return nil }
        let c = Int(dims[1]), cpp = c < 3 ? c : 4 * ((c + 3) / 4)
        var out = [Float](repeating: 0, count: Int(dims[0]) * Int(dims[2]) * Int(dims[3]) * cpp)
        let rc = out.withUnsafeMutableBufferPointer { smelter_tensor_to_float_mps(handle, nil, $0.baseAddress, $0.count) }
        return rc == 0 ? out : nil
    }
    /// Enqueue the read-back without waiting (read `into` in the stream's completion callback, like an MTLCommandBuffer handler).
    public func toFloatArrayAsync(into buffer: UnsafeMutableBufferPointer<Float>, stream: UnsafeMutableRawPointer? = nil) throws {
        try ONNXGraph.check(smelter_tensor_to_float_async(handle, stream, buffer.baseAddress, buffer.count))
    }
    /// MPSImage.toFloatArray() (MPSImage+Extensions.swift:9-59); NCHW order.
    public func toFloatArray() -> [Float]? {
        var dims = [Int32](repeating: 0, count: 4)
        guard smelter_tensor_dims(handle, &dims) == 0 else { return nil }
        var out = [Float](repeating: 0, count: dims.reduce(1) { $0 * Int($1) })
        let rc = out.withUnsafeMutableBufferPointer { smelter_tensor_to_float(handle, nil, $0.baseAddress, $0.count) }
        return rc == 0 ? out : nil
    }
    deinit { if owned { smelter_tensor_destroy(handle) } }
}

/// Replaces MPSNNGraph: `encode(to:sourceImages:)` enqueues on the stream (README.md:43-44).
public final class CUDANNGraph {
    let owner: ONNXGraph
    init(owner: ONNXGraph) { self.owner = owner }
    public func encode(to stream: UnsafeMutableRawPointer? = nil, sourceImages: [DeviceImage]) -> DeviceImage? {
        var sources: [OpaquePointer?] = sourceImages.map { Optional($0.handle) }
        var result: OpaquePointer?
        let rc = sources.withUnsafeMutableBufferPointer { smelter_graph_encode(owner.handle, stream, $0.baseAddress, Int32($0.count), &result) }
        guard rc == 0, let r = result else { return nil }
        return DeviceImage(borrowed: r)
    }
}

public final class ONNXGraph {
    /// ONNXGraph.swift:6-36
    public struct Configuration {
        public enum InputConstraint { case none, forceInputScale(ScaleAlgorithm) }
        public enum ScaleAlgorithm { case lanczos, bilinear }
        public struct BillinearUpsampling {
            public var alignCorners: Bool
            public static let `default` = BillinearUpsampling(alignCorners: true)
        }
        public var inputConstraint: InputConstraint
        public var billinearUpsamplingConfiguration: BillinearUpsampling
        public var dims: [Int: Int]
        /// Engine option (no reference counterpart, smelter_config.sm_share): k >= 2 sizes every kernel of the graph for 1/k of the
        /// SMs so that encodes in flight on different streams co-run; 1 = whole chip.
        public var smShare: Int32 = 1
        public init(inputConstraint: InputConstraint = .none, billinearUpsamplingConfiguration: BillinearUpsampling = .default, dims: [Int: Int] = [:]) {
            self.inputConstraint = inputConstraint
            self.billinearUpsamplingConfiguration = billinearUpsamplingConfiguration
            self.dims = dims
        }
    }

    /// ONNXGraph.swift:38-47 (+ engine codes >= 100 folded into graphInternalError)
    public enum Errors: Error {
        case unsupportedInput, unsupportedOutput, unknownNodeOpType(opType: String), noSuchOutput, graphInternalError
        case insufficientInputs, inconsistentState, notEnoughAttributes
    }
    /// ONNXGraph.swift:49-52
    public enum Format { case onnx, mpsFlavor }

    let handle: OpaquePointer
    public let configuration: Configuration
    private var compiled: CUDANNGraph?

    static func check(_ rc: Int32) throws {
        let message = String(cString: smelter_last_error())
        switch rc {
        case 0: return
        case 1: throw Errors.unsupportedInput
        case 2: throw Errors.unsupportedOutput
        case 3: throw Errors.unknownNodeOpType(opType: message)
        case 4: throw Errors.noSuchOutput
        case 6: throw Errors.insufficientInputs
        case 7: throw Errors.inconsistentState
        case 8: throw Errors.notEnoughAttributes
        default: throw Errors.graphInternalError
        }
    }

    /// ONNXGraph.init(data:configuration:) (ONNXGraph.swift:95); the context replaces the MTLDevice passed later in the reference.
    public init(data: Data, configuration: Configuration = .init(), context: CUDAContext) throws {
        self.configuration = configuration
        var cfg = smelter_config()
        smelter_config_default(&cfg)
        if case let .forceInputScale(alg) = configuration.inputConstraint { cfg.input_constraint = alg == .lanczos ? 1 : 2 }
        cfg.bilinear_align_corners = configuration.billinearUpsamplingConfiguration.alignCorners ? 1 : 0
        var n: Int32 = 0
        withUnsafeMutablePointer(to: &cfg.dims_axis) { axes in
            withUnsafeMutablePointer(to: &cfg.dims_value) { values in
                let a = UnsafeMutableRawPointer(axes).assumingMemoryBound(to: Int32.self)
                let v = UnsafeMutableRawPointer(values).assumingMemoryBound(to: Int64.self)
                for (axis, value) in configuration.dims.sorted(by: { $0.key < $1.key }).prefix(8) {
                    a[Int(n)] = Int32(axis); v[Int(n)] = Int64(value); n += 1
                }
            }
        }
        cfg.n_dims = n
        cfg.sm_share = configuration.smShare
        var h: OpaquePointer?
        try data.withUnsafeBytes { bytes in
            try ONNXGraph.check(smelter_graph_create(context.handle, bytes.bindMemory(to: UInt8.self).baseAddress, bytes.count, &cfg, &h))
        }
        handle = h!
    }

    /// ONNXGraph.swift:158-167
    public convenience init(contentsOf url: URL, configuration: Configuration = .init(), context: CUDAContext) throws {
        try self.init(data: try Data(contentsOf: url), configuration: configuration, context: context)
    }

    /// ONNXGraph.swift:58
    public var modelFormat: Format {
        var f: Int32 = 0
        smelter_graph_format(handle, &f)
        return f == 1 ? .mpsFlavor : .onnx
    }

    /// ONNXGraph.swift:69-91
    public var outputShapes: [Shape] {
        var n: Int32 = 0
        smelter_graph_num_outputs(handle, &n)
        return (0 ..< n).compactMap { i in
            var s = smelter_shape()
            guard smelter_graph_output_shape(handle, i, &s) == 0 else { return nil }
            return Shape(channels: Int(s.channels), width: Int(s.width), height: Int(s.height), depth: Int(s.depth))
        }
    }

    /// metalGraph(device:) (ONNXGraph.swift:169-193): walk the nodes through the converter registry and compile.
    public func metalGraph() throws -> CUDANNGraph {
        if let g = compiled { return g }
        try ONNXGraph.check(smelter_graph_build(handle))
        let g = CUDANNGraph(owner: self)
        compiled = g
        return g
    }

    /// register(name:converter:) (ONNXGraph.swift:253-257) for host-language NodeConverters (NodeConverter.swift:3-5):
    /// the closure receives the node index and calls the smelter_add_* builders.
    public func register(name: String, converter: @escaping smelter_converter_fn, user: UnsafeMutableRawPointer? = nil) throws {
        try ONNXGraph.check(smelter_graph_register_converter(handle, name, converter, user))
    }

    deinit { smelter_graph_destroy(handle) }
}
