namespace Tensor.B200

// P/Invoke binding of libdeepnet_b200.so (include/dn_tensor.h).
// NOT COMPILED IN THIS REPOSITORY'S IMAGE (no dotnet/fsharpc available); kept ABI-faithful to the header by
// construction: sequential layout, cdecl, int64 everywhere, bool marshalled as one byte (see tests/test_abi.py,
// which checks the same offsets from C and from the ctypes mirror).
// Loading follows the reference's own mechanism for native libraries (Tensor/Tensor/NativeLib.fs:138-202,
// NativeLibName.Packaged "deepnet_b200" -> runtimes/linux-x64/native/libdeepnet_b200.so), exactly how
// Tensor/Tensor/Host/HostBLAS.fs:113-157 binds MKL.

open System
open System.Runtime.InteropServices

/// dn_dtype — order is ABI.
type DnDType =
    | F32 = 0 | F64 = 1 | I8 = 2 | U8 = 3 | I16 = 4 | U16 = 5
    | I32 = 6 | U32 = 7 | I64 = 8 | U64 = 9 | Bool = 10

/// dn_status and the exception raised for it.
type DnStatus =
    | Ok = 0 | InvalidArg = 1 | Unsupported = 2 | OutOfMemory = 3 | IndexOutOfRange = 4
    | Cuda = 5 | NoDevice = 6 | ShapeMismatch = 7 | SingularMatrix = 8

/// dn_tensor — replaces NativeTensor (Tensor/Tensor/Cuda/NativeTensor.fs:50-57). 152 bytes.
[<Struct; StructLayout(LayoutKind.Sequential)>]
type DnTensor =
    val mutable Base:   nativeint
    val mutable Offset: int64
    val mutable NDims:  int32
    val mutable DType:  int32
    [<MarshalAs(UnmanagedType.ByValArray, SizeConst = 8)>]
    val mutable Shape:  int64[]
    [<MarshalAs(UnmanagedType.ByValArray, SizeConst = 8)>]
    val mutable Stride: int64[]

module Native =
    [<Literal>]
    let Lib = "deepnet_b200"

    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_init(int device)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_set_stream(nativeint stream)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_sync()
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_set_check_errors(int enabled)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern nativeint dn_last_error()
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_alloc(int64 nbytes, nativeint& ptr)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_free(nativeint ptr)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_memcpy_h2d(nativeint dstDev, nativeint srcHost, int64 nbytes)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_memcpy_d2h(nativeint dstHost, nativeint srcDev, int64 nbytes)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_get_item(DnTensor& t, int64[] pos, nativeint value)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_set_item(DnTensor& t, int64[] pos, nativeint value)

    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_fill_const(DnTensor& t, nativeint value)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_fill_incrementing(DnTensor& t, nativeint start, nativeint incr)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_copy(DnTensor& t, DnTensor& a)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_convert(DnTensor& t, DnTensor& a)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_unary(int op, DnTensor& t, DnTensor& a)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_binary(int op, DnTensor& t, DnTensor& a, DnTensor& b)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_compare(int op, DnTensor& t, DnTensor& a, DnTensor& b)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_is_finite(DnTensor& t, DnTensor& a)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_if_then_else(DnTensor& t, DnTensor& cond, DnTensor& ifTrue, DnTensor& ifFalse)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_reduce_last_axis(int op, DnTensor& t, DnTensor& a)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_arg_reduce_last_axis(int op, DnTensor& t, DnTensor& a)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_find_last_axis(nativeint value, DnTensor& t, DnTensor& a)
    // index / mask tensors: array of pointers to pinned DnTensor structs, IntPtr.Zero = None / NoMask
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_gather(DnTensor& t, nativeint[] idxs, int nidxs, DnTensor& a)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_scatter(DnTensor& t, nativeint[] idxs, int nidxs, DnTensor& a)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_count_true(DnTensor& a, int64& count)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_masked_get(DnTensor& t, DnTensor& a, nativeint[] masks, int nmasks)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_masked_set(DnTensor& t, nativeint[] masks, int nmasks, DnTensor& a)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_true_indices(DnTensor& t, DnTensor& a)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_vec_vec_dot(DnTensor& t, DnTensor& a, DnTensor& b)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_mat_vec_dot(DnTensor& t, DnTensor& a, DnTensor& b)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_mat_mat_dot(DnTensor& t, DnTensor& a, DnTensor& b)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_batched_mat_mat_dot(DnTensor& t, DnTensor& a, DnTensor& b)
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_batched_invert(DnTensor& t, DnTensor& a)
    /// dn_fused_instr: kind (0 unary, 1 binary, 2 const), op, dst, a, b, imm — see include/dn_tensor.h
    [<Struct; StructLayout(LayoutKind.Sequential)>]
    type DnFusedInstr =
        val mutable Kind: int; val mutable Op: int; val mutable Dst: int; val mutable A: int; val mutable B: int
        val mutable Imm: double
    [<DllImport(Lib, CallingConvention = CallingConvention.Cdecl)>]
    extern DnStatus dn_fused_elemwise(DnTensor& t, nativeint[] srcs, int nsrc, DnFusedInstr[] prog, int ninstr)

    /// Maps a non-OK status to the exception the reference raises in the same situation (SURVEY.md §8b).
    let check (st: DnStatus) =
        if st <> DnStatus.Ok then
            let msg = Marshal.PtrToStringAnsi (dn_last_error ())
            match st with
            | DnStatus.Unsupported      -> raise (NotSupportedException msg)
            | DnStatus.OutOfMemory      -> raise (OutOfMemoryException msg)   // OutOfCudaMemoryException in Tensor.Cuda
            | DnStatus.IndexOutOfRange  -> raise (IndexOutOfRangeException msg)
            | DnStatus.InvalidArg       -> raise (ArgumentException msg)
            | DnStatus.ShapeMismatch    -> raise (InvalidOperationException msg)
            | DnStatus.SingularMatrix   -> raise (SingularMatrixException msg)
            | _                         -> failwithf "CUDA error: %s" msg      // ManagedCuda.CudaException
