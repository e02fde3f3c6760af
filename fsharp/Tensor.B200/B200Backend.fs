namespace Tensor.B200

// TensorB200Device / TensorB200Storage<'T> / TensorB200Backend<'T>: the three interfaces of
// Tensor/Tensor/TensorBackend.fs:14-146 implemented over libdeepnet_b200.so. Drop-in for
// Tensor/Tensor/Cuda/CudaBackend.fs:51-108,120-492,497-507 — `Tensor<'T>` code written against CudaTensor.Dev runs
// unchanged against B200Tensor.Dev. NOT COMPILED HERE (no .NET toolchain in this image).

open System
open System.Runtime.InteropServices
open Tensor
open Tensor.Utils
open Tensor.Backend
open Tensor.Host

module internal Marshalling =
    let dtypeOf (t: Type) =
        match t with
        | t when t = typeof<single> -> DnDType.F32  | t when t = typeof<double> -> DnDType.F64
        | t when t = typeof<sbyte>  -> DnDType.I8   | t when t = typeof<byte>   -> DnDType.U8
        | t when t = typeof<int16>  -> DnDType.I16  | t when t = typeof<uint16> -> DnDType.U16
        | t when t = typeof<int32>  -> DnDType.I32  | t when t = typeof<uint32> -> DnDType.U32
        | t when t = typeof<int64>  -> DnDType.I64  | t when t = typeof<uint64> -> DnDType.U64
        | t when t = typeof<bool>   -> DnDType.Bool
        | t -> raise (NotSupportedException (sprintf "the type %s is not supported by the B200 backend" t.Name))

    /// NativeTensor successor: layout -> DnTensor (rank > 8 is rejected).
    let desc (basePtr: nativeint) (layout: TensorLayout) (dtype: DnDType) =
        if layout.NDims > 8 then raise (NotSupportedException "tensors of rank > 8 are not supported")
        let pad (xs: int64 list) = Array.append (Array.ofList xs) (Array.zeroCreate (8 - xs.Length))
        let mutable d = DnTensor()
        d.Base <- basePtr; d.Offset <- layout.Offset; d.NDims <- layout.NDims; d.DType <- int dtype
        d.Shape <- pad layout.Shape; d.Stride <- pad layout.Stride
        d

    /// Boxes a scalar of 'T into unmanaged memory for the duration of `f` (scalars cross the ABI by pointer).
    let withScalar (value: 'T) (f: nativeint -> unit) =
        let h = GCHandle.Alloc ((if typeof<'T> = typeof<bool> then box (if unbox<bool> (box value) then 1uy else 0uy)
                                 else box value), GCHandleType.Pinned)
        try f (h.AddrOfPinnedObject ()) finally h.Free ()

/// Index / mask tensor lists (Gather, Scatter, MaskedGet, MaskedSet): `const dn_tensor *const *` with NULL for
/// None / NoMask. Replaces the Reflection.Emit'ed NativeIdxTensors struct (NativeTensor.fs:157-208,
/// Kernels/GatherScatter.cuh:9-19): the descriptors are written to unmanaged memory for the duration of the call
/// (DnTensor carries by-value arrays, so it is marshalled, not pinned) and passed as an array of pointers.
module internal B200Index =
    let private descSize = Marshal.SizeOf typeof<DnTensor>

    /// Runs `f ptrs` with ptrs.[i] = pointer to a native copy of the i-th descriptor, IntPtr.Zero for None.
    let private withDescs (descs: DnTensor option list) (f: nativeint[] -> DnStatus) =
        let n = List.length descs
        let block = Marshal.AllocHGlobal (max 1 n * descSize)
        try
            let ptrs =
                descs |> List.mapi (fun i d ->
                    match d with
                    | Some d -> let p = block + nativeint (i * descSize)
                                Marshal.StructureToPtr (d, p, false)
                                p
                    | None -> IntPtr.Zero)
                |> Array.ofList
            Native.check (f (if n = 0 then [| IntPtr.Zero |] else ptrs))
        finally Marshal.FreeHGlobal block

    let gather (trgt: DnTensor) (idxs: DnTensor option list) (src: DnTensor) =
        let mutable t, a = trgt, src
        withDescs idxs (fun p -> Native.dn_gather (&t, p, List.length idxs, &a))
    let scatter (trgt: DnTensor) (idxs: DnTensor option list) (src: DnTensor) =
        let mutable t, a = trgt, src
        withDescs idxs (fun p -> Native.dn_scatter (&t, p, List.length idxs, &a))
    let maskedGet (trgt: DnTensor) (src: DnTensor) (masks: DnTensor option []) =
        let mutable t, a = trgt, src
        withDescs (List.ofArray masks) (fun p -> Native.dn_masked_get (&t, &a, p, masks.Length))
    let maskedSet (trgt: DnTensor) (masks: DnTensor option []) (src: DnTensor) =
        let mutable t, a = trgt, src
        withDescs (List.ofArray masks) (fun p -> Native.dn_masked_set (&t, p, masks.Length, &a))

/// Host <-> device Transfer (CudaBackend.fs:206-270). The managed array is pinned for the duration of the call and
/// registered with CUDA when possible (CudaRegMem.fs:114-146; DN_ERR_INVALID_ARG = the reference's
/// CannotCudaRegisterMemoryException: fall back to plain pinning). Any pair of layouts goes through in ONE native
/// call: dn_transfer_h2d / dn_transfer_d2h do the strided packing and the on-device layout change that the
/// reference does with `src.Copy(order=RowMajor)` on the host and a temporary device tensor + CopyFrom.
module internal B200Transfer =
    let private withHostMem (hs: ITensorHostStorage) (f: nativeint -> unit) =
        use pin = hs.Pin ()
        let registered = Native.dn_host_register (pin.Ptr, hs.DataSizeInBytes) = DnStatus.Ok
        try f pin.Ptr
        finally
            // the copy may still be in flight on the thread's stream: the array must stay pinned until it has left
            // (the reference defers the unpin with a stream callback, CudaBackend.fs:232,252)
            Native.check (Native.dn_sync ())
            if registered then Native.dn_host_unregister pin.Ptr |> ignore

    let hostToDevice (devDesc: DnTensor) (hs: ITensorHostStorage) (hostLayout: TensorLayout) (dtype: DnDType) =
        withHostMem hs (fun p ->
            let mutable d, h = devDesc, Marshalling.desc p hostLayout dtype
            Native.check (Native.dn_transfer_h2d (&d, &h)))

    let deviceToHost (hs: ITensorHostStorage) (hostLayout: TensorLayout) (devDesc: DnTensor) (dtype: DnDType) =
        withHostMem hs (fun p ->
            let mutable h, d = Marshalling.desc p hostLayout dtype, devDesc
            Native.check (Native.dn_transfer_d2h (&h, &d)))

type TensorB200Storage<'T when 'T: (new: unit -> 'T) and 'T: struct and 'T :> ValueType> (nElems: int64) =
    let nElems = max nElems 1L                                   // CudaBackend.fs:56-58
    let mutable ptr = 0n
    do Native.check (Native.dn_alloc (nElems * sizeof64<'T>, &ptr))
    member this.Ptr = ptr
    /// Explicit release from the thread that used the storage: stream-ordered (dn_free).
    member this.Dispose () =
        if ptr <> 0n then
            Native.dn_free ptr |> ignore
            ptr <- 0n
            GC.SuppressFinalize this
    /// The finalizer thread owns no stream and may not even have the storage's device current: the release is only
    /// QUEUED on the owning device (dn_free_deferred) and carried out by the next dn_alloc / dn_sync of a thread that
    /// works there, behind a fence over the device's streams (the reference: CudaBackend.fs:73-74 + the
    /// keep-alive events of CudaUtils.fs:122-177).
    override this.Finalize () = if ptr <> 0n then Native.dn_free_deferred ptr |> ignore
    interface IDisposable with
        member this.Dispose () = this.Dispose ()
    interface ITensorStorage<'T> with
        member this.Backend layout = TensorB200Backend<'T> (layout, this) :> ITensorBackend<_>
        member this.Dev = TensorB200Device.Instance :> ITensorDevice

and TensorB200Backend<'T when 'T: (new: unit -> 'T) and 'T: struct and 'T :> ValueType>
        (layout: TensorLayout, storage: TensorB200Storage<'T>) =
    static let dt = Marshalling.dtypeOf typeof<'T>
    static let d (t: ITensorFrontend<'U>) =
        let s = t.Storage :?> TensorB200Storage<'U>
        Marshalling.desc s.Ptr t.Layout (Marshalling.dtypeOf typeof<'U>)
    let unary op (trgt: ITensorFrontend<'T>) (a: ITensorFrontend<'T>) =
        let mutable t, a = d trgt, d a in Native.check (Native.dn_unary (op, &t, &a))
    let binary op (trgt: ITensorFrontend<'T>) (a: ITensorFrontend<'T>) (b: ITensorFrontend<'T>) =
        let mutable t, a, b = d trgt, d a, d b in Native.check (Native.dn_binary (op, &t, &a, &b))
    let compare op (trgt: ITensorFrontend<bool>) (a: ITensorFrontend<'T>) (b: ITensorFrontend<'T>) =
        let mutable t, a, b = d trgt, d a, d b in Native.check (Native.dn_compare (op, &t, &a, &b))
    let reduce op (trgt: ITensorFrontend<'R>) (a: ITensorFrontend<'S>) =
        let mutable t, a = d trgt, d a in Native.check (Native.dn_reduce_last_axis (op, &t, &a))
    interface ITensorBackend<'T> with
        member this.Item
            with get idx = let mutable t = Marshalling.desc storage.Ptr layout dt
                           let buf = [| Unchecked.defaultof<'T> |]
                           let h = GCHandle.Alloc (buf, GCHandleType.Pinned)
                           try Native.check (Native.dn_get_item (&t, idx, h.AddrOfPinnedObject ())); buf.[0] finally h.Free ()
            and set idx v = let mutable t = Marshalling.desc storage.Ptr layout dt
                            Marshalling.withScalar v (fun p -> Native.check (Native.dn_set_item (&t, idx, p)))
        member this.FillConst (value, trgt) =
            let mutable t = d trgt in Marshalling.withScalar value (fun p -> Native.check (Native.dn_fill_const (&t, p)))
        member this.FillIncrementing (start, incr, trgt) =
            let mutable t = d trgt
            Marshalling.withScalar start (fun s -> Marshalling.withScalar incr (fun i ->
                Native.check (Native.dn_fill_incrementing (&t, s, i))))
        member this.Copy (trgt, src) = let mutable t, a = d trgt, d src in Native.check (Native.dn_copy (&t, &a))
        member this.Convert (trgt, src) = let mutable t, a = d trgt, d src in Native.check (Native.dn_convert (&t, &a))
        member this.Transfer (trgt, src) =
            match trgt.Storage, src.Storage with
            | (:? TensorB200Storage<'T>), (:? TensorHostStorage<'T> as hs) ->
                B200Transfer.hostToDevice (d trgt) (hs :> ITensorHostStorage) src.Layout dt; true
            | (:? TensorHostStorage<'T> as hs), (:? TensorB200Storage<'T>) ->
                B200Transfer.deviceToHost (hs :> ITensorHostStorage) trgt.Layout (d src) dt; true
            | _ -> false
        // unary: op codes are dn_unary_op
        member this.UnaryPlus (t, a) = unary 0 t a
        member this.UnaryMinus (t, a) = unary 1 t a
        member this.Abs (t, a) = unary 2 t a
        member this.Sgn (t, a) = unary 3 t a
        member this.Log (t, a) = unary 4 t a
        member this.Log10 (t, a) = unary 5 t a
        member this.Exp (t, a) = unary 6 t a
        member this.Sin (t, a) = unary 7 t a
        member this.Cos (t, a) = unary 8 t a
        member this.Tan (t, a) = unary 9 t a
        member this.Asin (t, a) = unary 10 t a
        member this.Acos (t, a) = unary 11 t a
        member this.Atan (t, a) = unary 12 t a
        member this.Sinh (t, a) = unary 13 t a
        member this.Cosh (t, a) = unary 14 t a
        member this.Tanh (t, a) = unary 15 t a
        member this.Sqrt (t, a) = unary 16 t a
        member this.Ceiling (t, a) = unary 17 t a
        member this.Floor (t, a) = unary 18 t a
        member this.Round (t, a) = unary 19 t a
        member this.Truncate (t, a) = unary 20 t a
        member this.Negate (t, a) = let mutable t, a = d t, d a in Native.check (Native.dn_unary (21, &t, &a))
        member this.IsFinite (t, a) = let mutable t, a = d t, d a in Native.check (Native.dn_is_finite (&t, &a))
        // binary: dn_binary_op
        member this.Add (t, a, b) = binary 0 t a b
        member this.Subtract (t, a, b) = binary 1 t a b
        member this.Multiply (t, a, b) = binary 2 t a b
        member this.Divide (t, a, b) = binary 3 t a b
        member this.Modulo (t, a, b) = binary 4 t a b
        member this.Power (t, a, b) = binary 5 t a b
        member this.MaxElemwise (t, a, b) = binary 6 t a b
        member this.MinElemwise (t, a, b) = binary 7 t a b
        member this.And (t, a, b) = let mutable t, a, b = d t, d a, d b in Native.check (Native.dn_binary (8, &t, &a, &b))
        member this.Or (t, a, b) = let mutable t, a, b = d t, d a, d b in Native.check (Native.dn_binary (9, &t, &a, &b))
        member this.Xor (t, a, b) = let mutable t, a, b = d t, d a, d b in Native.check (Native.dn_binary (10, &t, &a, &b))
        // comparisons: dn_compare_op
        member this.Equal (t, a, b) = compare 0 t a b
        member this.NotEqual (t, a, b) = compare 1 t a b
        member this.Less (t, a, b) = compare 2 t a b
        member this.LessOrEqual (t, a, b) = compare 3 t a b
        member this.Greater (t, a, b) = compare 4 t a b
        member this.GreaterOrEqual (t, a, b) = compare 5 t a b
        member this.IfThenElse (t, c, a, b) =
            let mutable t, c, a, b = d t, d c, d a, d b in Native.check (Native.dn_if_then_else (&t, &c, &a, &b))
        // reductions: dn_reduce_op / dn_arg_reduce_op
        member this.SumLastAxis (t, a) = reduce 0 t a
        member this.ProductLastAxis (t, a) = reduce 1 t a
        member this.MinLastAxis (t, a) = reduce 2 t a
        member this.MaxLastAxis (t, a) = reduce 3 t a
        member this.AllLastAxis (t, a) = reduce 4 t a
        member this.AnyLastAxis (t, a) = reduce 5 t a
        member this.CountTrueLastAxis (t, a) = reduce 6 t a
        member this.ArgMinLastAxis (t, a) = let mutable t, a = d t, d a in Native.check (Native.dn_arg_reduce_last_axis (0, &t, &a))
        member this.ArgMaxLastAxis (t, a) = let mutable t, a = d t, d a in Native.check (Native.dn_arg_reduce_last_axis (1, &t, &a))
        member this.FindLastAxis (value, t, a) =
            let mutable t, a = d t, d a
            Marshalling.withScalar value (fun p -> Native.check (Native.dn_find_last_axis (p, &t, &a)))
        // indexing: option lists become arrays of pinned descriptor pointers, None -> IntPtr.Zero
        member this.Gather (t, idxs, a) = B200Index.gather (d t) (idxs |> List.map (Option.map d)) (d a)
        member this.Scatter (t, idxs, a) = B200Index.scatter (d t) (idxs |> List.map (Option.map d)) (d a)
        member this.MaskedGet (t, a, masks) = B200Index.maskedGet (d t) (d a) (masks |> Array.map (Option.map d))
        member this.MaskedSet (t, masks, a) = B200Index.maskedSet (d t) (masks |> Array.map (Option.map d)) (d a)
        member this.TrueIndices (t, a) = let mutable t, a = d t, d a in Native.check (Native.dn_true_indices (&t, &a))
        // dense contractions
        member this.VecVecDot (t, a, b) = let mutable t, a, b = d t, d a, d b in Native.check (Native.dn_vec_vec_dot (&t, &a, &b))
        member this.MatVecDot (t, a, b) = let mutable t, a, b = d t, d a, d b in Native.check (Native.dn_mat_vec_dot (&t, &a, &b))
        member this.MatMatDot (t, a, b) = let mutable t, a, b = d t, d a, d b in Native.check (Native.dn_mat_mat_dot (&t, &a, &b))
        member this.BatchedMatMatDot (t, a, b) =
            let mutable t, a, b = d t, d a, d b in Native.check (Native.dn_batched_mat_mat_dot (&t, &a, &b))
        // outside the hot path (SURVEY.md §8f-4); SVD / eig are unsupported in the reference's CUDA backend as well
        member this.BatchedInvert (t, a) =
            let mutable t, a = d t, d a in Native.check (Native.dn_batched_invert (&t, &a))
        member this.BatchedSVD (s, uv, a) = raise (NotSupportedException "BatchedSVD is not supported")
        member this.SymmetricEigenDecomposition (p, vals, vecs, a) =
            raise (NotSupportedException "SymmetricEigenDecomposition is not supported")

and TensorB200Device private () =
    inherit BaseTensorDevice ()
    static do Native.check (Native.dn_init 0)                    // CudaInit.check, CudaBackend.fs:28-38
    static member Instance = TensorB200Device ()
    override this.Id = "Cuda"                                     // same device id: existing code keeps working
    override this.Create nElems = TensorB200Storage<'T> nElems :> ITensorStorage<'T>
    override this.Zeroed = false

/// Frontend helpers: the analogue of module CudaTensor (Tensor/Tensor/Cuda/CudaFrontend.fs:34-150).
module B200Tensor =
    /// The B200 device; `Tensor<'T>` code written against CudaTensor.Dev runs unchanged against it.
    let Dev = TensorB200Device.Instance :> ITensorDevice
    let transfer x = Tensor.transfer Dev x
    let empty<'T> = Tensor<'T>.empty Dev
    let zeros<'T> = Tensor<'T>.zeros Dev
    let ones<'T> = Tensor<'T>.ones Dev
    let filled<'T> = Tensor<'T>.filled Dev
    let scalar<'T> = Tensor<'T>.scalar Dev
    let counting = Tensor.counting Dev
    /// Cfg.Stream (CudaCfg.fs:25-27): the calling thread's stream.
    let setStream (stream: nativeint) = Native.check (Native.dn_set_stream stream)
    /// Cfg.Stacktrace (CudaCfg.fs:33-35).
    let setStacktrace (enabled: bool) = Native.check (Native.dn_set_check_errors (if enabled then 1 else 0))
    let synchronize () = Native.check (Native.dn_sync ())
