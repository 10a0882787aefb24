{-# LANGUAGE CPP #-}
-- |
-- Module      : Streamly.Internal.LZ4.B200
-- Portability : GHC
--
-- Batched replacements for 'compressChunksD' / 'decompressChunksRawD' of
-- "Streamly.Internal.LZ4" (src/Streamly/Internal/LZ4.hs:353-394, :539-567)
-- over libb200lz4.so (include/b200lz4.h).  Same stream types, same block
-- header layout, same acceleration parameter, same error texts; the only
-- behavioural difference is that up to 'batchArrays' arrays (or 'batchBytes'
-- bytes) are pulled from upstream before one @safe@ FFI call is made, and the
-- results are then yielded one by one in order.
--
-- STATUS: written against streamly-0.8's internal Array / StreamD API exactly
-- as the reference uses it, but NOT type-checked in this repository's build
-- image (no GHC there).  Every behaviour it relies on is exercised at the C
-- ABI by tests/test_gpu_parity.py through streamly_lz4_b200/api.py, which is
-- the same logic in Python.  Drop this file next to
-- src/Streamly/Internal/LZ4.hs and re-export from there (INTEGRATION.md 2).
--
module Streamly.Internal.LZ4.B200
    ( compressChunksD
    , decompressChunksRawD
    , B200Config (..)
    , defaultB200Config
    )

where

import Control.Monad (forM, forM_, when)
import Control.Monad.IO.Class (MonadIO(..))
import Data.Int (Int32, Int64)
import Data.Word (Word8)
import Foreign.C (CInt(..), CSize(..), CString, peekCString)
import Foreign.Marshal.Alloc (alloca)
import Foreign.Marshal.Array (allocaArray, peekArray, pokeArray)
import Foreign.Marshal.Utils (copyBytes)
import Foreign.Ptr (Ptr, nullPtr, plusPtr)
import Foreign.Storable (peek, peekByteOff, poke)
import Fusion.Plugin.Types (Fuse (..))

import qualified Streamly.Internal.Data.Array.Foreign as Array
import qualified Streamly.Internal.Data.Array.Foreign.Type as Array
import qualified Streamly.Internal.Data.Array.Foreign.Mut.Type as MArray
import qualified Streamly.Internal.Data.Stream.StreamD as Stream

import Streamly.Internal.LZ4.Config

#define INLINE_NORMAL INLINE [1]
#define INLINE_LATE   INLINE [0]

--------------------------------------------------------------------------------
-- Foreign (include/b200lz4.h)
--------------------------------------------------------------------------------

data C_Ctx
data C_CStream
data C_DStream

foreign import ccall unsafe "b200lz4.h b200lz4_compress_bound"
    c_compressBound :: CInt -> IO CInt

foreign import ccall safe "b200lz4.h b200lz4_ctx_create"
    c_ctxCreate :: CInt -> Ptr (Ptr C_Ctx) -> IO CInt

foreign import ccall safe "b200lz4.h b200lz4_ctx_destroy"
    c_ctxDestroy :: Ptr C_Ctx -> IO ()

foreign import ccall safe "b200lz4.h b200lz4_host_alloc"
    c_hostAlloc :: CSize -> IO (Ptr Word8)

foreign import ccall safe "b200lz4.h b200lz4_host_free"
    c_hostFree :: Ptr Word8 -> IO ()

foreign import ccall unsafe "b200lz4.h b200lz4_last_error"
    c_lastError :: IO CString

foreign import ccall safe "b200lz4.h b200lz4_cstream_create"
    c_cstreamCreate :: Ptr C_Ctx -> Ptr (Ptr C_CStream) -> IO CInt

foreign import ccall safe "b200lz4.h b200lz4_cstream_free"
    c_cstreamFree :: Ptr C_CStream -> IO ()

foreign import ccall safe "b200lz4.h b200lz4_dstream_create"
    c_dstreamCreate :: Ptr C_Ctx -> Ptr (Ptr C_DStream) -> IO CInt

foreign import ccall safe "b200lz4.h b200lz4_dstream_free"
    c_dstreamFree :: Ptr C_DStream -> IO ()

foreign import ccall safe "b200lz4.h b200lz4_compress_batch"
    c_compressBatch
        :: Ptr C_Ctx
        -> Ptr Word8 -> Int64 -> Ptr Int64 -> Ptr Int32 -> CInt
        -> Ptr Int32 -> CInt -> Ptr (Ptr C_CStream)
        -> CInt -> CInt
        -> Ptr Word8 -> Int64 -> Ptr Int64 -> Ptr Int32
        -> IO CInt

foreign import ccall safe "b200lz4.h b200lz4_decompress_batch"
    c_decompressBatch
        :: Ptr C_Ctx
        -> Ptr Word8 -> Int64 -> Ptr Int64 -> Ptr Int32 -> CInt
        -> Ptr Int32 -> CInt -> Ptr (Ptr C_DStream)
        -> CInt -> CInt
        -> Ptr Word8 -> Int64 -> Ptr Int64 -> Ptr Int32
        -> IO CInt

lz4_MAX_INPUT_SIZE :: Int
lz4_MAX_INPUT_SIZE = 0x7E000000           -- B200LZ4_MAX_INPUT_SIZE

--------------------------------------------------------------------------------
-- Batching parameters
--------------------------------------------------------------------------------

data B200Config = B200Config
    { device :: Int            -- ^ CUDA device ordinal
    , batchArrays :: Int       -- ^ at most this many arrays per FFI call
    , batchBytes :: Int        -- ^ ... or this many input bytes
    }

defaultB200Config :: B200Config
defaultB200Config = B200Config 0 4096 (256 * 1024 * 1024)

-- | One library context plus two growable page-locked staging buffers.
data Session = Session
    { sCtx :: Ptr C_Ctx
    , sSrc :: Ptr Word8, sSrcCap :: Int
    , sDst :: Ptr Word8, sDstCap :: Int
    }

lastError :: IO String
lastError = c_lastError >>= peekCString

newSession :: B200Config -> IO Session
newSession conf = alloca $ \pp -> do
    rc <- c_ctxCreate (fromIntegral (device conf)) pp
    when (rc /= 0) $ lastError >>= \e -> error ("b200lz4_ctx_create failed: " ++ e)
    ctx <- peek pp
    let cap = batchBytes conf + 16 * (batchArrays conf + 1)
    src <- c_hostAlloc (fromIntegral cap)
    return $ Session ctx src cap nullPtr 0

freeSession :: Session -> IO ()
freeSession s = do
    c_hostFree (sSrc s)
    when (sDst s /= nullPtr) $ c_hostFree (sDst s)
    c_ctxDestroy (sCtx s)

ensureDst :: Session -> Int -> IO Session
ensureDst s need
    | need <= sDstCap s = return s
    | otherwise = do
        when (sDst s /= nullPtr) $ c_hostFree (sDst s)
        p <- c_hostAlloc (fromIntegral (need + need `div` 4))
        return s { sDst = p, sDstCap = need + need `div` 4 }

ensureSrc :: Session -> Int -> IO Session
ensureSrc s need
    | need <= sSrcCap s = return s
    | otherwise = do
        c_hostFree (sSrc s)
        p <- c_hostAlloc (fromIntegral (need + need `div` 4))
        return s { sSrc = p, sSrcCap = need + need `div` 4 }

stagedSize :: [Array.Array Word8] -> Int
stagedSize = sum . map (\a -> align16 (Array.byteLength a + 16))

align16 :: Int -> Int
align16 n = (n + 15) `div` 16 * 16

-- | Copy the arrays into the pinned source buffer at 16-byte aligned offsets
-- with a 16-byte gap (separate Haskell arrays are never adjacent) and return
-- (offsets, lengths).
stage :: Session -> [Array.Array Word8] -> IO ([Int64], [Int32])
stage s arrs = go 0 arrs [] []
  where
    go _ [] offs lens = return (reverse offs, reverse lens)
    go at (a:as) offs lens = do
        let n = Array.byteLength a
        Array.asPtrUnsafe (Array.unsafeCast a) $ \p ->
            copyBytes (sSrc s `plusPtr` at) (p :: Ptr Word8) n
        go (at + align16 (n + 16)) as (fromIntegral at : offs) (fromIntegral n : lens)

-- | Slice a fresh, exactly sized array out of the pinned destination buffer.
sliceOut :: Ptr Word8 -> Int -> Int -> IO (Array.Array Word8)
sliceOut base off len = do
    (MArray.Array cont b0 b dmax) <- MArray.newArray len
    copyBytes b (base `plusPtr` off) len
    return $ Array.unsafeFreeze (MArray.Array cont b0 (b `plusPtr` len) dmax)

--------------------------------------------------------------------------------
-- One batch through the library
--------------------------------------------------------------------------------

maxBlockSizeOf :: BlockConfig -> Int
maxBlockSizeOf cfg =
    case blockSize cfg of
        BlockHasSize -> lz4_MAX_INPUT_SIZE
        BlockMax64KB -> 64 * 1024
        BlockMax256KB -> 256 * 1024
        BlockMax1MB -> 1024 * 1024
        BlockMax4MB -> 4 * 1024 * 1024

-- | 'compressChunk' (src/Streamly/Internal/LZ4.hs:226-281) for a whole batch.
compressBatch
    :: BlockConfig -> Int -> Session -> Ptr C_CStream -> [Array.Array Word8]
    -> IO (Session, [Array.Array Word8])
compressBatch cfg speed s0 strm arrs = do
    let n = length arrs
        meta = metaSize cfg
    forM_ arrs $ \a -> do
        let len = Array.byteLength a
        when (len >= 2 * 1024 * 1024 * 1024)
            $ error "compressChunksD: Array element > 2 GB encountered"
        when (len > maxBlockSizeOf cfg)
            $ error $ "compressChunk: Source array length " ++ show len
                ++ " exceeds the maximum block size of " ++ show (maxBlockSizeOf cfg)
    bounds <- forM arrs $ \a -> do
        b <- c_compressBound (fromIntegral (Array.byteLength a))
        when (b <= 0) $ error "compressChunk: compressed length <= 0."
        return (fromIntegral b + meta)
    s <- ensureDst s0 (sum bounds) >>= \s' -> ensureSrc s' (stagedSize arrs)
    (offs, lens) <- stage s arrs
    let srcBytes = case (offs, lens) of
            ([], _) -> 0
            _ -> fromIntegral (last offs) + fromIntegral (last lens)
    allocaArray n $ \pOff -> allocaArray n $ \pLen ->
      allocaArray (n + 1) $ \pDstOff -> allocaArray n $ \pOutLen ->
      allocaArray 2 $ \pFirst -> alloca $ \pStrm -> do
        pokeArray pOff offs
        pokeArray pLen lens
        pokeArray pFirst [0, fromIntegral n]
        poke pStrm strm
        -- linked mode: all blocks of the batch go through the stream's one state, in order
        -- (one LZ4_stream_t per Haskell stream, :367-376); independent: no stream table
        let (pF, nS, pS) = if strm == nullPtr then (nullPtr, 0, nullPtr) else (pFirst, 1, pStrm)
        rc <- c_compressBatch (sCtx s) (sSrc s) srcBytes pOff pLen (fromIntegral n)
                  pF nS pS (fromIntegral speed) (fromIntegral meta)
                  (sDst s) (fromIntegral (sDstCap s)) pDstOff pOutLen
        when (rc /= 0) $ lastError >>= \e ->
            error ("compressChunk: c_compressFastContinue failed. " ++ e)
        dstOff <- peekArray (n + 1) pDstOff
        outs <- forM (zip dstOff (tail dstOff)) $ \(a, b) ->
            sliceOut (sDst s) (fromIntegral a) (fromIntegral (b - a))
        return (s, outs)

-- | 'decompressChunk' (src/Streamly/Internal/LZ4.hs:290-336) for a whole batch of framed arrays.
decompressBatch
    :: BlockConfig -> Session -> Ptr C_DStream -> [Array.Array Word8]
    -> IO (Session, [Array.Array Word8])
decompressBatch cfg s0 strm arrs = do
    let n = length arrs
        meta = metaSize cfg
        maxBlock = case blockSize cfg of
            BlockHasSize -> 0
            _ -> maxBlockSizeOf cfg
    caps <- forM arrs $ \a ->
        if meta == 8 && Array.byteLength a >= 8
        then Array.asPtrUnsafe (Array.unsafeCast a) $ \p -> do
                 u <- peekByteOff (p :: Ptr Word8) 4 :: IO Int32     -- uncompLen LE32 (little-endian host)
                 return (max 0 (fromIntegral u))
        else return maxBlock
    s <- ensureDst s0 (sum caps + 64) >>= \s' -> ensureSrc s' (stagedSize arrs)
    (offs, lens) <- stage s arrs
    let srcBytes = case (offs, lens) of
            ([], _) -> 0
            _ -> fromIntegral (last offs) + fromIntegral (last lens)
    allocaArray n $ \pOff -> allocaArray n $ \pLen ->
      allocaArray (n + 1) $ \pDstOff -> allocaArray n $ \pOutLen ->
      allocaArray 2 $ \pFirst -> alloca $ \pStrm -> do
        pokeArray pOff offs
        pokeArray pLen lens
        pokeArray pFirst [0, fromIntegral n]
        poke pStrm strm
        let (pF, nS, pS) = if strm == nullPtr then (nullPtr, 0, nullPtr) else (pFirst, 1, pStrm)
        rc <- c_decompressBatch (sCtx s) (sSrc s) srcBytes pOff pLen (fromIntegral n)
                  pF nS pS (fromIntegral meta) (fromIntegral maxBlock)
                  (sDst s) (fromIntegral (sDstCap s)) pDstOff pOutLen
        when (rc /= 0) $ lastError >>= \e ->
            error ("decompressChunk: c_decompressSafeContinue failed. " ++ e)
        dstOff <- peekArray n pDstOff
        outLen <- peekArray n pOutLen
        outs <- forM (zip dstOff outLen) $ \(a, l) ->
            sliceOut (sDst s) (fromIntegral a) (fromIntegral l)
        return (s, outs)

--------------------------------------------------------------------------------
-- Stream combinators
--------------------------------------------------------------------------------

{-# ANN type BatchState Fuse #-}
data BatchState st ses
    = BInit st
    | BFill st ses [Array.Array Word8] Int Int     -- pending arrays (reversed), count, bytes
    | BDrain st ses [Array.Array Word8] Bool       -- ready outputs; True = upstream finished
    | BDone ses

-- | Shared driver: pull up to a batch, run it, yield the results in order.
{-# INLINE_NORMAL batchedD #-}
batchedD ::
       MonadIO m
    => B200Config
    -> IO ses                                              -- ^ acquire (context, stream state)
    -> (ses -> IO ())                                      -- ^ release
    -> (ses -> [Array.Array Word8] -> IO (ses, [Array.Array Word8]))
    -> Stream.Stream m (Array.Array Word8)
    -> Stream.Stream m (Array.Array Word8)
batchedD conf acquire release run (Stream.Stream step0 state0) =
    Stream.Stream step (BInit state0)

    where

    flush st ses pending finished = do
        (ses1, outs) <- liftIO $ run ses (reverse pending)
        return $ Stream.Skip $ BDrain st ses1 outs finished

    {-# INLINE_LATE step #-}
    step _ (BInit st) = do
        ses <- liftIO acquire
        return $ Stream.Skip $ BFill st ses [] 0 0
    step gst (BFill st ses pending cnt bytes)
        | cnt >= batchArrays conf || bytes >= batchBytes conf = flush st ses pending False
        | otherwise = do
            r <- step0 gst st
            case r of
                Stream.Yield arr st1 ->
                    return $ Stream.Skip
                        $ BFill st1 ses (arr : pending) (cnt + 1) (bytes + Array.byteLength arr)
                Stream.Skip st1 -> return $ Stream.Skip $ BFill st1 ses pending cnt bytes
                Stream.Stop ->
                    if null pending
                    then return $ Stream.Skip $ BDone ses
                    else flush st ses pending True
    step _ (BDrain st ses (a : as) fin) = return $ Stream.Yield a (BDrain st ses as fin)
    step _ (BDrain st ses [] False) = return $ Stream.Skip $ BFill st ses [] 0 0
    step _ (BDrain _ ses [] True) = return $ Stream.Skip $ BDone ses
    step _ (BDone ses) = liftIO (release ses) >> return Stream.Stop

-- | Drop-in for 'Streamly.Internal.LZ4.compressChunksD' (:353-394).  Linked
-- blocks by default (one device-resident stream state per Haskell stream);
-- independent blocks would pass 'nullPtr' as the stream (the reference's
-- 'setBlockIndependence' is still a stub, Config.hs:142-146).
{-# INLINE_NORMAL compressChunksD #-}
compressChunksD ::
       MonadIO m
    => B200Config
    -> BlockConfig
    -> Int
    -> Stream.Stream m (Array.Array Word8)
    -> Stream.Stream m (Array.Array Word8)
compressChunksD conf cfg speed0 =
    batchedD conf acquire release run

    where

    speed = max speed0 0                                        -- :364
    acquire = do
        s <- newSession conf
        strm <- alloca $ \pp -> do
            rc <- c_cstreamCreate (sCtx s) pp
            when (rc /= 0) $ lastError >>= \e -> error ("b200lz4_cstream_create failed: " ++ e)
            peek pp
        return (s, strm)
    release (s, strm) = c_cstreamFree strm >> freeSession s      -- :393-394
    run (s, strm) arrs = do
        (s1, outs) <- compressBatch cfg speed s strm arrs
        return ((s1, strm), outs)

-- | Drop-in for 'Streamly.Internal.LZ4.decompressChunksRawD' (:539-567): every
-- input array is exactly one framed block (what 'resizeChunksD' yields).
{-# INLINE_NORMAL decompressChunksRawD #-}
decompressChunksRawD ::
       MonadIO m
    => B200Config
    -> BlockConfig
    -> Stream.Stream m (Array.Array Word8)
    -> Stream.Stream m (Array.Array Word8)
decompressChunksRawD conf cfg =
    batchedD conf acquire release run

    where

    acquire = do
        s <- newSession conf
        strm <- alloca $ \pp -> do
            rc <- c_dstreamCreate (sCtx s) pp
            when (rc /= 0) $ lastError >>= \e -> error ("b200lz4_dstream_create failed: " ++ e)
            peek pp
        return (s, strm)
    release (s, strm) = c_dstreamFree strm >> freeSession s
    run (s, strm) arrs = do
        (s1, outs) <- decompressBatch cfg s strm arrs
        return ((s1, strm), outs)
